"""One corpus-mode query on a small shard (ncu target)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rag_docvqa_b200 import sharded
dev = torch.device("cuda:0")
N, d, Q, k = int(os.environ.get("N", 400000)), 768, int(os.environ.get("Q", 1024)), 10
g = torch.Generator(device=dev); g.manual_seed(1)
E = torch.randn(N, d, generator=g, device=dev).to(torch.bfloat16)
shard = sharded.CorpusShard(E)
Qf = torch.randn(Q, d, generator=g, device=dev)
for _ in range(3):
    shard.candidates(Qf, k)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); shard.candidates(Qf, k); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print("N=%d Q=%d: %.3f ms %.1f TFLOP/s" % (N, Q, ms, 2.0 * Q * N * d / ms / 1e9))
