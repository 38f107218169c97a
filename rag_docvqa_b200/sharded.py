"""Multi-GPU retrieval.  Per-document retrieval (C1-C4) shards by document with no collective: partition_documents /
take_documents below.  Corpus-scale retrieval: chunk embeddings row-sharded across the GPUs of one box
(BASELINE.json configs[4]: 10 M chunks x 768-d bf16 over 8 x B200, 1024 questions, top-10).

No reference counterpart (the reference is single-GPU and scores one question per document); the
scoring formula is Retriever._get_similarities' (src/_modules.py:1990-1993) applied to every
(question, chunk) pair.  One process per GPU:

  rank r owns rows [r*N/W, (r+1)*N/W) as bf16 + fp32 inverse norms (CorpusShard);
  search():  local tcgen05 score + fused top-k  ->  (Q, k) candidates with GLOBAL row ids
             -> ONE all-gather of (Q, k) values + ids over NCCL / NVLink (80 KB per rank at Q=1024, k=10)
             -> merge kernel by (score desc, global id asc) on every rank.
The same packed ordering is used locally and in the merge, so the sharded answer equals the
unsharded one bit for bit (tests/test_sharded_gloo.py checks the exchange logic on CPU ranks with
world_size 2, injecting the oracle's merge; tests/test_tc_gpu.py checks the kernels on the GPU).
"""
from __future__ import annotations

import json
import os
from typing import Optional, Tuple

import numpy as np
import torch

from . import _lib
from . import functional as F

_fn = _lib.lib


def shard_bounds(n_rows: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous row range of `rank` (balanced: sizes differ by at most one)."""
    base, rem = divmod(n_rows, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def partition_documents(sizes, world: int):
    """Per-document retrieval across ranks: every (question, document) pair is
    independent (src/_modules.py:1986-1995 loops over the batch with no cross-talk), so a batch is split by DOCUMENT,
    with no data-path collective.  Greedy balance on the chunk counts (largest document first onto the least loaded
    rank, ties to the lower rank), documents of a rank kept in batch order.  Deterministic from `sizes` alone, so every
    rank computes the same partition without talking.  Returns `world` lists of document indices."""
    if world < 1:
        raise ValueError("world must be >= 1")
    sizes = [int(x) for x in sizes]
    load = [0] * world
    parts = [[] for _ in range(world)]
    for b in sorted(range(len(sizes)), key=lambda i: (-sizes[i], i)):
        r = min(range(world), key=lambda j: (load[j], len(parts[j]), j))
        parts[r].append(b)
        load[r] += sizes[b]
    return [sorted(p) for p in parts]


def take_documents(indices, *per_document):
    """The rank's slice of any per-document arguments of Retriever.retrieve (lists, or a (B, d) question tensor)."""
    out = []
    for x in per_document:
        if isinstance(x, torch.Tensor):
            out.append(x[torch.as_tensor(indices, dtype=torch.long, device=x.device)] if len(indices) else x[:0])
        else:
            out.append([x[i] for i in indices])
    return out[0] if len(out) == 1 else tuple(out)


class CorpusShard:
    """bf16 rows + fp32 inverse norms of one rank's slice of the corpus, resident in HBM."""

    def __init__(self, rows_bf16: torch.Tensor, id_offset: int = 0, inv_norm: Optional[torch.Tensor] = None):
        if not rows_bf16.is_cuda:
            raise RuntimeError("CorpusShard lives on a CUDA device: rag_docvqa_b200 has no CPU fallback")
        if rows_bf16.dtype != torch.bfloat16 or rows_bf16.dim() != 2 or not rows_bf16.is_contiguous():
            raise ValueError("CorpusShard: expected a contiguous (n, d) bf16 matrix")
        self.rows = rows_bf16
        self.id_offset = int(id_offset)
        self.n, self.d = rows_bf16.shape
        if inv_norm is None:                                  # a rank may own no rows (N < world): nothing to norm
            inv_norm = F.bf16_inv_norm(rows_bf16) if self.n else torch.empty((0,), dtype=torch.float32, device=rows_bf16.device)
        self.inv_norm = inv_norm

    @classmethod
    def from_f32(cls, rows_f32: torch.Tensor, id_offset: int = 0) -> "CorpusShard":
        rows, inv = F.rows_to_bf16(rows_f32, normalise=False, return_inv_norm=True)
        return cls(rows, id_offset, inv)

    def candidates(self, questions: torch.Tensor, k: int):
        """Local scoring: (Q, groups*16) candidate values / global ids (id -1 = empty)."""
        dev = self.rows.device
        Qn = questions.shape[0]
        if self.n == 0:                                       # an empty shard still answers: no candidates (id -1)
            per = int(_fn.rdv_tc_candidates_per_group())
            return (torch.full((Qn, per), float("-inf"), dtype=torch.float32, device=dev),
                    torch.full((Qn, per), -1, dtype=torch.int64, device=dev))
        q_bf16, q_inv = F.rows_to_bf16(questions, normalise=False, return_inv_norm=True)
        groups = int(_fn.rdv_corpus_groups(self.n, Qn))
        tile_m, per = int(_fn.rdv_tc_tile_m()), int(_fn.rdv_tc_candidates_per_group())
        rows_padded = (Qn + tile_m - 1) // tile_m * tile_m
        part_val = torch.empty((groups, rows_padded, per), dtype=torch.float32, device=dev)
        part_idx = torch.empty((groups, rows_padded, per), dtype=torch.int32, device=dev)
        cand_val = torch.empty((Qn, groups * per), dtype=torch.float32, device=dev)
        cand_idx = torch.empty((Qn, groups * per), dtype=torch.int64, device=dev)
        with torch.cuda.device(dev):
            _lib.check(_fn.rdv_corpus_score_topk_bf16(
                self.rows.data_ptr(), self.inv_norm.data_ptr(), self.n, self.d, q_bf16.data_ptr(), q_inv.data_ptr(),
                Qn, k, self.id_offset, groups, part_val.data_ptr(), part_idx.data_ptr(), cand_val.data_ptr(),
                cand_idx.data_ptr(), F._stream_ptr(dev)))
        return cand_val, cand_idx

    def search_local(self, questions: torch.Tensor, k: int):
        """(Q, k) best local chunks: values (cosine) and GLOBAL ids, (score desc, id asc)."""
        cand_val, cand_idx = self.candidates(questions, k)
        return F.topk_merge(cand_val, cand_idx, k)


# ---- on-disk corpus index ---------------------------------------------------------
# A directory with  rows.bf16  (N x d bfloat16, row-major, raw),  inv_norm.f32  (N fp32 = 1 / ||row||, raw)  and
# meta.json.  Every rank memory-maps the files and uploads only ITS row range (shard_bounds), through a pinned
# staging buffer in fixed-size pieces, so a 15 GB index never exists twice in host memory.
INDEX_VERSION = 1


def save_corpus_index(path: str, rows_bf16: torch.Tensor, inv_norm: Optional[torch.Tensor] = None, chunk_rows: int = 1 << 18) -> None:
    """Writes a corpus (CPU or CUDA tensor, (N, d) bf16) as an index directory.  `inv_norm` is computed on the device
    for CUDA inputs when not given; for CPU inputs it is left to load time."""
    if rows_bf16.dtype != torch.bfloat16 or rows_bf16.dim() != 2:
        raise ValueError("save_corpus_index: expected an (N, d) bf16 matrix")
    os.makedirs(path, exist_ok=True)
    n, d = rows_bf16.shape
    if inv_norm is None and rows_bf16.is_cuda:
        inv_norm = F.bf16_inv_norm(rows_bf16.contiguous())
    with open(os.path.join(path, "rows.bf16"), "wb") as f:
        for a in range(0, n, chunk_rows):
            f.write(rows_bf16[a:a + chunk_rows].contiguous().cpu().view(torch.int16).numpy().tobytes())
    if inv_norm is not None:
        with open(os.path.join(path, "inv_norm.f32"), "wb") as f:
            f.write(inv_norm.detach().float().cpu().numpy().tobytes())
    with open(os.path.join(path, "meta.json"), "w") as f:
        json.dump({"version": INDEX_VERSION, "rows": int(n), "dim": int(d), "dtype": "bfloat16",
                   "has_inv_norm": inv_norm is not None}, f)


class CorpusIndex:
    """Read side of an index directory: memory-mapped rows, per-rank loading."""

    def __init__(self, path: str):
        with open(os.path.join(path, "meta.json")) as f:
            self.meta = json.load(f)
        if self.meta.get("version") != INDEX_VERSION or self.meta.get("dtype") != "bfloat16":
            raise ValueError("unsupported corpus index: %r" % (self.meta,))
        self.path, self.n, self.d = path, int(self.meta["rows"]), int(self.meta["dim"])
        self.rows = np.memmap(os.path.join(path, "rows.bf16"), dtype=np.int16, mode="r", shape=(self.n, self.d)) if self.n else \
            np.zeros((0, self.d), np.int16)
        self.inv_norm = (np.memmap(os.path.join(path, "inv_norm.f32"), dtype=np.float32, mode="r", shape=(self.n,))
                         if self.meta.get("has_inv_norm") and self.n else None)

    def load_shard(self, device, rank: int = 0, world: int = 1, chunk_rows: int = 1 << 18) -> CorpusShard:
        """Rows [lo, hi) of this rank on `device`, global ids starting at lo."""
        lo, hi = shard_bounds(self.n, world, rank)
        device = torch.device(device)
        with torch.cuda.device(device):                        # copies, events and the norm kernel all on `device`
            rows = torch.empty((hi - lo, self.d), dtype=torch.bfloat16, device=device)
            stage = [torch.empty((min(chunk_rows, max(hi - lo, 1)), self.d), dtype=torch.int16, pin_memory=True) for _ in range(2)]
            done = [None, None]
            stream = torch.cuda.current_stream(device)
            for i, a in enumerate(range(lo, hi, chunk_rows)):
                b = min(hi, a + chunk_rows)
                buf = stage[i & 1]
                if done[i & 1] is not None:
                    done[i & 1].synchronize()                  # the previous copy out of this staging buffer
                buf[:b - a].numpy()[...] = self.rows[a:b]
                rows[a - lo:b - lo].view(torch.int16).copy_(buf[:b - a], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(stream)                              # on the stream the copy was enqueued on
                done[i & 1] = ev
            inv = None
            if self.inv_norm is not None:
                inv = torch.from_numpy(np.ascontiguousarray(self.inv_norm[lo:hi])).to(device)
            torch.cuda.synchronize(device)
            return CorpusShard(rows, id_offset=lo, inv_norm=inv)


def merge_across_ranks(local_val: torch.Tensor, local_idx: torch.Tensor, k: int, group=None, merge_fn=None):
    """All-gather every rank's (Q, k) candidates and merge on every rank.  With world_size 1 (or no
    process group) the local result is already final.  `merge_fn(cand_val, cand_idx, k)` defaults to the
    CUDA merge kernel (rdv_topk_merge); there is no host implementation in the product -- the CPU (gloo)
    test of this exchange logic injects the oracle's merge."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local_val, local_idx
    world = dist.get_world_size(group)
    Q, kk = local_val.shape
    if local_val.is_cuda and (Q * kk) % 2 == 0:
        # ONE collective: values (fp32) and ids (int64) travel in one byte buffer of Q*k*12 bytes per rank
        nv = Q * kk * 4
        send = torch.empty(nv + Q * kk * 8, dtype=torch.uint8, device=local_val.device)
        send[:nv].view(torch.float32).copy_(local_val.reshape(-1))
        send[nv:].view(torch.int64).copy_(local_idx.reshape(-1))
        recv = torch.empty((world, send.numel()), dtype=torch.uint8, device=local_val.device)
        dist.all_gather_into_tensor(recv, send, group=group)
        cand_val = recv[:, :nv].contiguous().view(torch.float32).view(world, Q, kk).permute(1, 0, 2).reshape(Q, world * kk)
        cand_idx = recv[:, nv:].contiguous().view(torch.int64).view(world, Q, kk).permute(1, 0, 2).reshape(Q, world * kk)
    else:
        vals = [torch.empty_like(local_val) for _ in range(world)]
        idxs = [torch.empty_like(local_idx) for _ in range(world)]
        dist.all_gather(vals, local_val.contiguous(), group=group)
        dist.all_gather(idxs, local_idx.contiguous(), group=group)
        cand_val, cand_idx = torch.cat(vals, dim=1), torch.cat(idxs, dim=1)
    return (merge_fn or merge_candidates)(cand_val, cand_idx, k)


class CorpusSearcher:
    """The corpus-mode query for a fixed (Q, k) with every buffer allocated once and the whole step -- bf16 cast of the
    questions, tcgen05 score + fused top-k on the local shard (rdv_corpus_score_topk_bf16), local merge written STRAIGHT
    into the NCCL send buffer, all-gather, final merge reading the receive buffer in NCCL's own rank-major layout
    (rdv_topk_merge_parts) -- captured into ONE CUDA graph and replayed (`graph=True`).  Same kernels and ordering as
    search(): sharded == unsharded bit for bit.

    search(questions) copies the (Q, d) fp32 questions into the static input and replays; the returned (Q, k) value / id
    tensors are the searcher's own output buffers (valid until the next call)."""

    def __init__(self, shard: CorpusShard, n_questions: int, k: int, group=None, graph: bool = True):
        import torch.distributed as dist
        self.shard, self.Q, self.k, self.group = shard, int(n_questions), int(k), group
        dev = self.dev = shard.rows.device
        self.world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
        self._dist = dist if self.world > 1 else None
        Q, d = self.Q, shard.d
        f32, i32, i64 = dict(dtype=torch.float32, device=dev), dict(dtype=torch.int32, device=dev), dict(dtype=torch.int64, device=dev)
        self.q_in = torch.zeros((Q, d), **f32)
        self.q_bf16 = torch.empty((Q, d), dtype=torch.bfloat16, device=dev)
        self.q_inv = torch.empty((Q,), **f32)
        self.groups = int(_fn.rdv_corpus_groups(shard.n, Q)) if shard.n else 0
        tile_m, per = int(_fn.rdv_tc_tile_m()), int(_fn.rdv_tc_candidates_per_group())
        rows_padded = (Q + tile_m - 1) // tile_m * tile_m
        self.part_val = torch.empty((max(self.groups, 1), rows_padded, per), **f32)
        self.part_idx = torch.empty((max(self.groups, 1), rows_padded, per), **i32)
        self.cand_val = torch.empty((Q, max(self.groups, 1) * per), **f32)
        self.cand_idx = torch.empty((Q, max(self.groups, 1) * per), **i64)
        # one byte buffer per rank: (Q, k) fp32 values | pad to 8 | (Q, k) int64 global ids
        self.nv = (Q * k * 4 + 7) // 8 * 8
        self.send = torch.zeros(self.nv + Q * k * 8, dtype=torch.uint8, device=dev)
        self.send_val = self.send[:Q * k * 4].view(torch.float32).view(Q, k)
        self.send_idx = self.send[self.nv:].view(torch.int64).view(Q, k)
        if shard.n == 0:                                      # no rows: the (constant) contribution is "no candidates"
            self.send_val.fill_(float("-inf"))
            self.send_idx.fill_(-1)
        if self.world > 1:
            self.recv = torch.empty((self.world, self.send.numel()), dtype=torch.uint8, device=dev)
            self.out_val = torch.empty((Q, k), **f32)
            self.out_idx = torch.empty((Q, k), **i64)
        else:
            self.recv, self.out_val, self.out_idx = None, self.send_val, self.send_idx
        self.graphed = False
        self._graph = None
        if graph:
            self._capture()

    # -- the three parts of a step, each enqueued on torch's current stream ------------------------------------------
    def _local(self):
        sh, s = self.shard, F._stream_ptr(self.dev)
        if sh.n == 0:
            return
        _lib.check(_fn.rdv_rows_to_bf16(self.q_in.data_ptr(), self.Q, sh.d, 0, self.q_bf16.data_ptr(), self.q_inv.data_ptr(), s))
        _lib.check(_fn.rdv_corpus_score_topk_bf16(
            sh.rows.data_ptr(), sh.inv_norm.data_ptr(), sh.n, sh.d, self.q_bf16.data_ptr(), self.q_inv.data_ptr(), self.Q,
            self.k, sh.id_offset, self.groups, self.part_val.data_ptr(), self.part_idx.data_ptr(), self.cand_val.data_ptr(),
            self.cand_idx.data_ptr(), s))
        _lib.check(_fn.rdv_topk_merge(self.cand_val.data_ptr(), self.cand_idx.data_ptr(), self.Q, self.cand_val.shape[1],
                                      self.k, self.send_val.data_ptr(), self.send_idx.data_ptr(), s))

    def _exchange(self):
        if self.world > 1:
            self._dist.all_gather_into_tensor(self.recv, self.send, group=self.group)

    def _final(self):
        if self.world > 1:
            stride = self.send.numel()
            _lib.check(_fn.rdv_topk_merge_parts(self.recv.data_ptr(), self.recv.data_ptr() + self.nv, self.Q, self.world, self.k,
                                                stride // 4, stride // 8, self.k, self.out_val.data_ptr(), self.out_idx.data_ptr(),
                                                F._stream_ptr(self.dev)))

    def _step(self):
        self._local()
        self._exchange()
        self._final()

    def _capture(self):
        with torch.cuda.device(self.dev):
            side = torch.cuda.Stream(self.dev)
            side.wait_stream(torch.cuda.current_stream(self.dev))
            with torch.cuda.stream(side):
                for _ in range(2):                            # NCCL communicator + kernel attributes set up before capture
                    self._step()
            torch.cuda.current_stream(self.dev).wait_stream(side)
            torch.cuda.synchronize(self.dev)
            g = torch.cuda.CUDAGraph()
            try:
                with torch.cuda.graph(g, stream=side):
                    self._step()
                self._graph, self.graphed = g, True
            except Exception:                                  # capture not possible here (e.g. a collective that refuses it)
                self._graph, self.graphed = None, False
                torch.cuda.synchronize(self.dev)

    def search(self, questions: torch.Tensor):
        if tuple(questions.shape) != tuple(self.q_in.shape):
            raise ValueError("CorpusSearcher: expected (%d, %d) questions, got %s" % (self.Q, self.shard.d, tuple(questions.shape)))
        with torch.cuda.device(self.dev):
            self.q_in.copy_(questions, non_blocking=True)
            if self._graph is not None:
                self._graph.replay()
            else:
                self._step()
        return self.out_val, self.out_idx

    def close(self):
        """Releases the captured step.  Call it (or drop the searcher) BEFORE torch.distributed.destroy_process_group():
        a live CUDA graph that holds NCCL work keeps the communicator busy and the teardown waits for it forever."""
        if self._graph is not None:
            torch.cuda.synchronize(self.dev)
            self._graph = None
            self.graphed = False

    # -- parts alone, for measurement ---------------------------------------------------------------------------------
    def local_only(self, questions: torch.Tensor):
        with torch.cuda.device(self.dev):
            self.q_in.copy_(questions, non_blocking=True)
            self._local()
        return self.send_val, self.send_idx

    def exchange_only(self):
        with torch.cuda.device(self.dev):
            self._exchange()
            self._final()


def merge_candidates(cand_val: torch.Tensor, cand_idx: torch.Tensor, k: int):
    """(Q, m) -> (Q, k) by (score desc, id asc) on the device (rdv_topk_merge).  CUDA tensors only."""
    return F.topk_merge(cand_val, cand_idx, k)


def search(shard: CorpusShard, questions: torch.Tensor, k: int, group=None):
    """The corpus-mode query: local tcgen05 scoring + top-k, all-gather, merge."""
    val, idx = shard.search_local(questions, k)
    return merge_across_ranks(val, idx, k, group=group)
