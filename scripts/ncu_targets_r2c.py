"""Two launches of the corpus kernel at the per-rank shard of N = 8 (1.25 M rows x 1024 questions x 768-d bf16), where
the launcher picks CTA pairs (tc_score_kernel<2>).

    ncu --set full --clock-control none -k regex:tc_score -o /tmp/r2c python scripts/ncu_targets_r2c.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from rag_docvqa_b200 import sharded

dev = torch.device("cuda:0")
torch.cuda.set_device(0)
g = torch.Generator(device=dev).manual_seed(5)
rows = torch.randn(1_250_000, 768, generator=g, device=dev, dtype=torch.float32).to(torch.bfloat16)
shard = sharded.CorpusShard(rows, id_offset=0)
q = torch.randn(1024, 768, generator=g, device=dev)
for _ in range(2):
    val, idx = shard.search_local(q, 10)
torch.cuda.synchronize()
print(idx[0].tolist())
