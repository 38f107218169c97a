"""Two launches each of the kernels added late in round 2: the cluster-split pooling kernel and the pooled-patch selection
(C4p shape), and the generator-input embedding kernel (C2's packed shape).

    ncu --set full --clock-control none -k regex:'mean_pool_split|pooled_doc|topk_segments|vt5_embed' -o /tmp/r2b \
        python scripts/ncu_targets_r2b.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from rag_docvqa_b200 import functional as F, synth

dev = torch.device("cuda:0")
torch.cuda.set_device(0)
patches, q = synth.make_strip_batch(8, [50] * 8, 2048, 768, 1234, device=dev)
for _ in range(2):
    res = F.pooled_patch_topk(patches, q, 5)
torch.cuda.synchronize()
print("pooled:", res.patch_idx[0].tolist())
del patches, q, res
torch.cuda.empty_cache()


class Ctx:
    pass


Ctx.dev = dev
out = bench.embed_leg(Ctx, 64, 512, 6523.7, with_cpu=False)
print({k: v for k, v in out.items() if k != "what"})
