"""One launch each of the rank-3 / rank-4 kernels at C2 shapes (ncu target; see scripts/final_gpu_run.sh)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from rag_docvqa_b200 import postproc, synth
from rag_docvqa_b200.chunker import Chunker

dev = torch.device("cuda:0")
rng = np.random.RandomState(0)
B, k, n_b = 64, 5, 600
scores = torch.from_numpy(rng.rand(B, k).astype(np.float32)).to(dev)
cnt = torch.full((B,), k, dtype=torch.int32, device=dev)
pages = torch.from_numpy(rng.randint(0, 20, size=(B, k)).astype(np.int32)).to(dev)
sims = torch.from_numpy(rng.rand(B * n_b).astype(np.float32)).to(dev)
row_off = torch.arange(0, (B + 1) * n_b, n_b, dtype=torch.int64, device=dev)
postproc.rerank_order(scores, cnt, 0.4, 5, 1)
postproc.page_vote(pages, cnt, None, row_off, False)
postproc.page_vote(pages, cnt, sims, row_off, True)
words, boxes, info = synth.make_chunker_batch(77, docs=64, max_pages=20, max_words=700, max_layouts=30, degenerate=False)
ch = Chunker({"compute_stats": False, "compute_stats_examples": False, "n_stats_examples": 0, "device": "cuda:0"})
pb = [np.asarray(p, dtype=np.float64).reshape(-1, 4) for d in boxes for p in d]
lb = [np.asarray(pg["boxes"], dtype=np.float64).reshape(-1, 4) for d in info for pg in d]
ch.assign_words_to_layouts(pb, lb, [np.arange(len(x), dtype=np.int32) for x in lb])
torch.cuda.synchronize()
print("ok")
