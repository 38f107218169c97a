"""Quick timing probe of the tensor-core kernels (not the bench of record)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rag_docvqa_b200 import functional as F, sharded, _lib

dev = torch.device("cuda:0")
def timeit(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

N, d, Q, k = 1_250_000, 768, 1024, 10
g = torch.Generator(device=dev); g.manual_seed(1)
E = torch.randn(N, d, generator=g, device=dev, dtype=torch.float32).to(torch.bfloat16)
shard = sharded.CorpusShard(E)
Qf = torch.randn(Q, d, generator=g, device=dev)
ms = timeit(lambda: shard.candidates(Qf, k))
print("corpus candidates: %.3f ms  %.1f TFLOP/s" % (ms, 2.0 * Q * N * d / ms / 1e9))
ms2 = timeit(lambda: shard.search_local(Qf, k))
print("corpus search_local (incl. merge): %.3f ms  %.0f q/s" % (ms2, Q / ms2 * 1e3))
for Qn in (128, 256, 512):
    q2 = Qf[:Qn].contiguous()
    ms = timeit(lambda: shard.candidates(q2, k))
    print("  Q=%d: %.3f ms %.1f TFLOP/s" % (Qn, ms, 2.0 * Qn * N * d / ms / 1e9))
del E, shard
p = torch.randn(50, 2048, 768, generator=g, device=dev)
q = torch.randn(1, 2048, 768, generator=g, device=dev)
ms = timeit(lambda: F.late_interaction_bf16(q, p))
print("maxsim bf16 tc (incl. normalise+cast): %.3f ms  %.1f TFLOP/s" % (ms, 2.0 * 50 * 2048 * 2048 * 768 / ms / 1e9))
qn = F.rows_to_bf16(q[0], normalise=True); pn = F.rows_to_bf16(p, normalise=True)
part = torch.empty(50 * 16, device=dev); out = torch.empty(50, device=dev)
s = torch.cuda.current_stream().cuda_stream
ms = timeit(lambda: _lib.lib.rdv_maxsim_bf16_tc(qn.data_ptr(), pn.data_ptr(), 50, 2048, 2048, 768, part.data_ptr(), out.data_ptr(), s))
print("maxsim bf16 tc kernel only: %.3f ms  %.1f TFLOP/s" % (ms, 2.0 * 50 * 2048 * 2048 * 768 / ms / 1e9))
ms = timeit(lambda: F.late_interaction(q, p), 3)
print("maxsim f32: %.3f ms %.1f TFLOP/s" % (ms, 2.0 * 50 * 2048 * 2048 * 768 / ms / 1e9))
