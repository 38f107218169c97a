"""One C4-shaped MaxSim question through the tf32x3 kernel (ncu target)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rag_docvqa_b200 import functional as F, synth
dev = torch.device("cuda:0")
patches, q = synth.make_strip_batch(1, [50], 2048, 768, 3, device=dev)
for _ in range(3):
    out = F.late_interaction(q[0:1], patches[0], mode="tf32x3")
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); out = F.late_interaction(q[0:1], patches[0], mode="tf32x3"); e1.record(); torch.cuda.synchronize()
print("maxsim tf32x3: %.3f ms" % e0.elapsed_time(e1))
