"""Corpus mode over REAL NCCL (one process per GPU, world_size = min(4, visible GPUs)): the row-sharded search --
local tcgen05 score + top-k, all-gather, merge over the receive layout, all inside one captured CUDA graph
(sharded.CorpusSearcher) and through the eager path (sharded.search) -- equals the unsharded search bit for bit.
Needs >= 2 GPUs; skipped on a single-GPU box (there the exchange logic is covered by tests/test_sharded_gloo.py and the
kernels by tests/test_tc_gpu.py)."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
sys.path.insert(0, %r)
import torch
import torch.distributed as dist
from rag_docvqa_b200 import sharded
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
ok = True
def say(*a):
    print("[rank %%d]" %% rank, *a, flush=True)
# scenario 1: rows sharded evenly; scenario 2: rank 0 owns every row, the other ranks own EMPTY shards (N < world in miniature)
for scenario, (N, d, Q, k) in enumerate(((200_000, 256, 300, 10), (5_000, 64, 9, 5))):
    g = torch.Generator(device="cpu").manual_seed(5)
    E = (torch.randn(N, d, generator=g) + 0.3).to(torch.bfloat16)
    E[N - 7] = E[11]                                                 # an exact tie (across shards in scenario 1)
    lo, hi = sharded.shard_bounds(N, world, rank) if scenario == 0 else ((0, N) if rank == 0 else (N, N))
    shard = sharded.CorpusShard(E[lo:hi].to(dev).contiguous(), id_offset=lo)
    whole = sharded.CorpusShard(E.to(dev).contiguous())
    say("scenario", scenario, "rows", lo, hi)
    searcher = sharded.CorpusSearcher(shard, Q, k, graph=True)
    say("searcher built, graphed =", searcher.graphed)
    for seed in (1, 2):
        Qs = torch.randn(Q, d, generator=torch.Generator(device="cpu").manual_seed(seed)).to(dev)
        ref_v, ref_i = whole.search_local(Qs, k)
        v1, i1 = sharded.search(shard, Qs, k)
        v2, i2 = searcher.search(Qs)
        torch.cuda.synchronize()
        say("seed", seed, "done")
        ok = ok and torch.equal(i1, ref_i) and torch.equal(v1, ref_v) and torch.equal(i2, ref_i) and torch.equal(v2, ref_v)
    ok = ok and (searcher.graphed or world == 1)
    searcher.close()          # a captured graph that holds NCCL work must be released before the process group goes away
    del searcher
flag = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("sharded == unsharded on %%d ranks: %%s" %% (world, bool(flag.item())), flush=True)
torch.cuda.synchronize()
dist.destroy_process_group()
sys.exit(0 if flag.item() else 1)
'''


def test_sharded_search_over_nccl_equals_unsharded(tmp_path):
    n_gpus = torch.cuda.device_count()
    if n_gpus < 2:
        pytest.skip("needs >= 2 GPUs (gpurun --gpus 2)")
    world = min(4, n_gpus)
    script = tmp_path / "worker.py"
    script.write_text(WORKER % ROOT)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr",
           "127.0.0.1", "--master-port", str(port), str(script)]
    try:
        res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=240,
                             env={**os.environ, "NCCL_DEBUG": "WARN"})
    except subprocess.TimeoutExpired as exc:                          # a hang: show how far the ranks got
        out = exc.stdout.decode("utf-8", "replace") if isinstance(exc.stdout, bytes) else (exc.stdout or "")
        pytest.fail("the NCCL workers did not finish in 240 s; their output so far:\n" + out[-4000:])
    assert res.returncode == 0, res.stdout[-3000:]
    assert "sharded == unsharded on %d ranks: True" % world in res.stdout
