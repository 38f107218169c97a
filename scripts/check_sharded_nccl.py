"""torchrun check: the NCCL-sharded corpus search equals the unsharded search bit for bit (run on >= 2 GPUs)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from rag_docvqa_b200 import sharded
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
g = torch.Generator(device="cpu").manual_seed(5)
N, d, Q, k = 200_000, 256, 300, 10
E = (torch.randn(N, d, generator=g) + 0.3).to(torch.bfloat16)
E[N - 7] = E[11]                                          # a cross-shard exact tie
Qs = torch.randn(Q, d, generator=g).to(dev)
lo, hi = sharded.shard_bounds(N, world, rank)
shard = sharded.CorpusShard(E[lo:hi].to(dev).contiguous(), id_offset=lo)
val, idx = sharded.search(shard, Qs, k)
whole = sharded.CorpusShard(E.to(dev).contiguous())
ref_v, ref_i = whole.search_local(Qs, k)
ok = torch.equal(idx, ref_i) and torch.equal(val, ref_v)
flag = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("sharded == unsharded on %d ranks: %s" % (world, bool(flag.item())))
dist.destroy_process_group()
sys.exit(0 if flag.item() else 1)
