"""Where a C4p step's time goes: host time per call (no sync inside the loop), device time per step, and a cProfile of the host side."""
import cProfile
import pstats
import sys
import time

import torch

sys.path.insert(0, ".")
from rag_docvqa_b200 import functional as F, synth  # noqa: E402

dev = torch.device("cuda:0")
B, strips, L, d, k = 8, 50, 2048, 768, 5
patches, q = synth.make_strip_batch(B, [strips] * B, L, d, 1234, device=dev)
for _ in range(5):
    F.pooled_patch_topk(patches, q, k)
torch.cuda.synchronize()
n = 50
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
e0.record()
for _ in range(n):
    F.pooled_patch_topk(patches, q, k)
e1.record()
t_host = (time.perf_counter() - t0) / n
torch.cuda.synchronize()
print("host issue time / step: %.1f us   device time / step: %.1f us" % (t_host * 1e6, e0.elapsed_time(e1) / n * 1e3))
# host only, device kept busy by a long kernel queue is the same thing; profile it
pr = cProfile.Profile()
pr.enable()
for _ in range(n):
    F.pooled_patch_topk(patches, q, k)
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
