"""Corpus-scale retrieval: chunk embeddings row-sharded across the GPUs of one box
(BASELINE.json configs[4]: 10 M chunks x 768-d bf16 over 8 x B200, 1024 questions, top-10).

No reference counterpart (the reference is single-GPU and scores one question per document); the
scoring formula is Retriever._get_similarities' (src/_modules.py:1990-1993) applied to every
(question, chunk) pair.  One process per GPU:

  rank r owns rows [r*N/W, (r+1)*N/W) as bf16 + fp32 inverse norms (CorpusShard);
  search():  local tcgen05 score + fused top-k  ->  (Q, k) candidates with GLOBAL row ids
             -> ONE all-gather of (Q, k) values + ids over NCCL / NVLink (80 KB per rank at Q=1024, k=10)
             -> merge kernel by (score desc, global id asc) on every rank.
The same packed ordering is used locally and in the merge, so the sharded answer equals the
unsharded one bit for bit (tests/test_sharded_gloo.py checks this on CPU with world_size 2 for the
host logic; tests/test_tc_gpu.py on the GPU).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _lib
from . import functional as F

_fn = _lib.lib


def shard_bounds(n_rows: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous row range of `rank` (balanced: sizes differ by at most one)."""
    base, rem = divmod(n_rows, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class CorpusShard:
    """bf16 rows + fp32 inverse norms of one rank's slice of the corpus, resident in HBM."""

    def __init__(self, rows_bf16: torch.Tensor, id_offset: int = 0, inv_norm: Optional[torch.Tensor] = None):
        if not rows_bf16.is_cuda:
            raise RuntimeError("CorpusShard lives on a CUDA device: rag_docvqa_b200 has no CPU fallback")
        if rows_bf16.dtype != torch.bfloat16 or rows_bf16.dim() != 2 or not rows_bf16.is_contiguous():
            raise ValueError("CorpusShard: expected a contiguous (n, d) bf16 matrix")
        self.rows = rows_bf16
        self.id_offset = int(id_offset)
        self.inv_norm = inv_norm if inv_norm is not None else F.bf16_inv_norm(rows_bf16)
        self.n, self.d = rows_bf16.shape

    @classmethod
    def from_f32(cls, rows_f32: torch.Tensor, id_offset: int = 0) -> "CorpusShard":
        rows, inv = F.rows_to_bf16(rows_f32, normalise=False, return_inv_norm=True)
        return cls(rows, id_offset, inv)

    def candidates(self, questions: torch.Tensor, k: int):
        """Local scoring: (Q, groups*16) candidate values / global ids (id -1 = empty)."""
        dev = self.rows.device
        Qn = questions.shape[0]
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


def merge_across_ranks(local_val: torch.Tensor, local_idx: torch.Tensor, k: int, group=None):
    """All-gather every rank's (Q, k) candidates and merge on every rank.  With world_size 1 (or no
    process group) the local result is already final."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local_val, local_idx
    world = dist.get_world_size(group)
    vals = [torch.empty_like(local_val) for _ in range(world)]
    idxs = [torch.empty_like(local_idx) for _ in range(world)]
    dist.all_gather(vals, local_val.contiguous(), group=group)
    dist.all_gather(idxs, local_idx.contiguous(), group=group)
    cand_val, cand_idx = torch.cat(vals, dim=1), torch.cat(idxs, dim=1)
    return merge_candidates(cand_val, cand_idx, k)


def merge_candidates(cand_val: torch.Tensor, cand_idx: torch.Tensor, k: int):
    """(Q, m) -> (Q, k) by (score desc, id asc).  CUDA tensors go through rdv_topk_merge; the torch
    path below exists ONLY so the host-side sharding logic can be exercised with the gloo backend on
    CPU ranks in tests (it is never taken on a GPU rank)."""
    if cand_val.is_cuda:
        return F.topk_merge(cand_val, cand_idx, k)
    return _merge_candidates_host(cand_val, cand_idx, k)


def _merge_candidates_host(cand_val, cand_idx, k):
    big = torch.iinfo(torch.int64).max
    val = torch.where(cand_idx >= 0, cand_val, torch.full_like(cand_val, float("-inf")))
    idx = torch.where(cand_idx >= 0, cand_idx, torch.full_like(cand_idx, big))
    order = torch.argsort(idx, dim=1, stable=True)                      # id asc ...
    val, idx = torch.gather(val, 1, order), torch.gather(idx, 1, order)
    order = torch.argsort(val, dim=1, descending=True, stable=True)     # ... then score desc (stable)
    val, idx = torch.gather(val, 1, order)[:, :k], torch.gather(idx, 1, order)[:, :k]
    idx = torch.where(idx == big, torch.full_like(idx, -1), idx)
    return val, idx


def search(shard: CorpusShard, questions: torch.Tensor, k: int, group=None):
    """The corpus-mode query: local tcgen05 scoring + top-k, all-gather, merge."""
    val, idx = shard.search_local(questions, k)
    return merge_across_ranks(val, idx, k, group=group)
