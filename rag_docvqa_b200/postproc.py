"""What consumes the retrieved hits before generation: the reranker's
post-processing and the page vote, on the device.

Mirrors, with the reference's names and argument meaning,
  * `Reranker.rerank` / `batch_rerank` (src/_modules.py:1541-1610): the cross-encoder is a model and out of scope --
    it is passed in (`cross_encoder.forward(pairs) -> scores`, the reference's own optional constructor argument,
    :1544-1556); everything after its call (argsort, threshold, max / min clamp, permutation) runs through
    `rdv_rerank_order`;
  * the `majorpage` / `weightmajorpage` vote of `RAGVT5.forward` (src/RAGVT5.py:455-475): `major_page_indices`.
On the packed path (`Retriever.retrieve_packed`) nothing returns to Python lists: `rerank_packed` hands the index
list to the gather kernel, which rebuilds `input_ids / boxes / mask` in the reranked order.
"""
from __future__ import annotations

import ctypes
from typing import Any, List, Optional, Sequence

import numpy as np
import torch

from . import _lib
from .functional import _require_cuda, _stream_ptr

MAX_K = 64


def rerank_order(scores: torch.Tensor, cnt: Optional[torch.Tensor] = None, filter_thresh: float = 0.4,
                 max_chunk_num: int = 5, min_chunk_num: int = 1):
    """scores: (B, k) float32 / float64 CUDA tensor of cross-encoder scores in retrieve()'s output order; cnt: (B,)
    int32 candidates per document (None = k).  Returns (order (B,k) int32, -1 padded; kept (B,) int32; the scores in
    the new order (B,k)).  One launch, no synchronisation."""
    _require_cuda(scores, "rerank_order: scores")
    if scores.dim() != 2 or scores.dtype not in (torch.float32, torch.float64):
        raise ValueError("rerank_order: scores must be (B, k) float32 or float64")
    scores = scores.contiguous()
    B, k = scores.shape
    dev = scores.device
    if cnt is not None:
        _require_cuda(cnt, "rerank_order: cnt")
        if cnt.dtype != torch.int32 or tuple(cnt.shape) != (B,):
            raise ValueError("rerank_order: cnt must be (B,) int32")
        cnt = cnt.contiguous()
    order = torch.empty((B, k), dtype=torch.int32, device=dev)
    kept = torch.empty((B,), dtype=torch.int32, device=dev)
    out_scores = torch.empty_like(scores)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib.rdv_rerank_order(scores.data_ptr(), 1 if scores.dtype == torch.float64 else 0,
                                             cnt.data_ptr() if cnt is not None else None, B, k, float(filter_thresh),
                                             int(max_chunk_num), int(min_chunk_num), order.data_ptr(), kept.data_ptr(),
                                             out_scores.data_ptr(), _stream_ptr(dev)))
    return order, kept, out_scores


def page_vote(hit_page: torch.Tensor, hit_cnt: torch.Tensor, sims: Optional[torch.Tensor], row_off: torch.Tensor,
              weighted: bool, legacy_promotion: bool = True, return_weight: bool = False):
    """hit_page (B,k) int32 / hit_cnt (B,) int32: top_k_page_indices; sims (N,) float32 + row_off (B+1,) int64: every
    similarity of the batch.  Returns major (B,) int32 (and the winning weights, float64)."""
    for t, what in ((hit_page, "hit_page"), (hit_cnt, "hit_cnt"), (row_off, "row_off")):
        _require_cuda(t, "page_vote: " + what)
    if hit_page.dtype != torch.int32 or hit_cnt.dtype != torch.int32 or row_off.dtype != torch.int64:
        raise ValueError("page_vote: hit_page / hit_cnt int32, row_off int64")
    B, k = hit_page.shape
    if tuple(hit_cnt.shape) != (B,) or tuple(row_off.shape) != (B + 1,):
        raise ValueError("page_vote: hit_cnt (B,), row_off (B+1,)")
    dev = hit_page.device
    if sims is not None:
        _require_cuda(sims, "page_vote: sims")
        if sims.dtype != torch.float32:
            raise ValueError("page_vote: sims must be float32")
        sims = sims.contiguous()
    elif weighted:
        raise ValueError("page_vote: weightmajorpage needs the similarities")
    major = torch.empty((B,), dtype=torch.int32, device=dev)
    weight = torch.empty((B,), dtype=torch.float64, device=dev) if return_weight else None
    hit_page, hit_cnt, row_off = hit_page.contiguous(), hit_cnt.contiguous(), row_off.contiguous()
    with torch.cuda.device(dev):
        _lib.check(_lib.lib.rdv_page_vote(hit_page.data_ptr(), hit_cnt.data_ptr(),
                                          sims.data_ptr() if sims is not None else None, row_off.data_ptr(), B, k,
                                          1 if weighted else 0, 1 if legacy_promotion else 0, major.data_ptr(),
                                          weight.data_ptr() if weight is not None else None, _stream_ptr(dev)))
    return (major, weight) if return_weight else major


def _pad_rows(rows: Sequence[Sequence[float]], k: int, dtype) -> np.ndarray:
    out = np.zeros((len(rows), k), dtype=dtype)
    for b, r in enumerate(rows):
        out[b, :len(r)] = r
    return out


def major_page_indices(top_k_page_indices: Sequence[Sequence[int]], similarities: Sequence[torch.Tensor],
                       page_retrieval: str = "majorpage", device=None, legacy_promotion: bool = True) -> List[int]:
    """`major_page_indices` of RAGVT5.forward (src/RAGVT5.py:455-475) from retrieve()'s own outputs:
    top_k_page_indices [B][k] and similarities List[B] Tensor(n_b,)."""
    if page_retrieval not in ("majorpage", "weightmajorpage"):
        raise ValueError("page_retrieval must be 'majorpage' or 'weightmajorpage'")
    B = len(top_k_page_indices)
    if B == 0:
        return []
    k = max(1, max(len(p) for p in top_k_page_indices))
    if k > MAX_K:
        raise _lib.RdvError(_lib.E_LIMIT, "page vote: more than %d hits per document" % MAX_K)
    dev = torch.device(device) if device is not None else next(
        (s.device for s in similarities if s.is_cuda), torch.device("cuda", torch.cuda.current_device()))
    sizes = [int(s.shape[0]) for s in similarities]
    row_off = np.zeros(B + 1, dtype=np.int64)
    np.cumsum(sizes, out=row_off[1:])
    pages = _pad_rows(top_k_page_indices, k, np.int32)
    cnt = np.asarray([len(p) for p in top_k_page_indices], dtype=np.int32)
    sims = torch.cat([s.to(dev, torch.float32).reshape(-1) for s in similarities]) if sum(sizes) else \
        torch.zeros((1,), dtype=torch.float32, device=dev)
    major = page_vote(torch.from_numpy(pages).to(dev), torch.from_numpy(cnt).to(dev), sims,
                      torch.from_numpy(row_off).to(dev), page_retrieval == "weightmajorpage", legacy_promotion)
    return [int(x) for x in major.cpu().tolist()]


class Reranker:
    """Drop-in for src._modules.Reranker (:1541-1610) given a cross-encoder object (`forward(pairs) -> scores`)."""

    def __init__(self, config: dict, cross_encoder: Optional[Any] = None):
        self.rerank_filter_tresh = float(config.get("rerank_filter_tresh", 0.4))
        self.rerank_max_chunk_num = config.get("rerank_max_chunk_num", 5)
        self.rerank_min_chunk_num = config.get("rerank_min_chunk_num", 1)
        if cross_encoder is None:
            raise ValueError("Reranker: pass the cross-encoder (src._modules.CrossEncoder / FlagLLMReranker); the "
                             "models are outside this package")
        self.cross_encoder = cross_encoder
        dev = config.get("device", "cuda")
        self.device = torch.device("cuda", torch.cuda.current_device()) if dev == "cuda" else torch.device(dev)

    def _scores(self, question: str, candidates: Sequence[str]):
        pairs = list(zip([question] * len(candidates), candidates))
        with torch.no_grad():
            return self.cross_encoder.forward(pairs)

    def _orders(self, score_rows) -> List[List[int]]:
        """One launch + one read for the whole batch."""
        B = len(score_rows)
        k = max(1, max((len(r) for r in score_rows), default=1))
        if k > MAX_K:
            raise _lib.RdvError(_lib.E_LIMIT, "rerank: more than %d candidates per document" % MAX_K)
        # CrossEncoder.predict returns a float32 array, FlagLLMReranker.compute_score a list of Python floats (float64
        # once np.argsort sees it); the kernel sorts whichever type arrives
        host = [r.detach().cpu().numpy() if isinstance(r, torch.Tensor) else np.asarray(r) for r in score_rows]
        host = [h if h.dtype in (np.float32, np.float64) else h.astype(np.float64) for h in host]
        dt = np.float64 if any(h.dtype == np.float64 and h.size for h in host) else np.float32
        scores = torch.from_numpy(_pad_rows(host, k, dt)).to(self.device)
        cnt = torch.tensor([len(r) for r in host], dtype=torch.int32).to(self.device)
        order, kept, _ = rerank_order(scores, cnt, self.rerank_filter_tresh, self.rerank_max_chunk_num,
                                      self.rerank_min_chunk_num)
        order_h, kept_h = order.cpu().numpy(), kept.cpu().numpy()
        return [order_h[b, :kept_h[b]].tolist() for b in range(B)]

    def rerank(self, question: str, candidates: List[str], *args: List[Any]) -> tuple:
        order = self._orders([self._scores(question, candidates)])[0]
        return ([candidates[i] for i in order], *[[arg[i] for i in order] for arg in args])

    def batch_rerank(self, questions: List[str], candidates: List[List[str]], *args: List[List[Any]]) -> tuple:
        rows = [self._scores(q, c) for q, c in zip(questions, candidates)]
        orders = self._orders(rows) if rows else []
        sorted_candidates = [[c[i] for i in o] for c, o in zip(candidates, orders)]
        sorted_args = [[[arg_b[i] for i in o] for arg_b, o in zip(arg, orders)] for arg in args]
        return (sorted_candidates, *sorted_args)

    def rerank_packed(self, plan, scores: torch.Tensor, pages=None, **pack_options):
        """Packed path: `plan` is the GatherPlan of Retriever.retrieve_packed(..., return_plan=True); `scores` (B,k)
        are the cross-encoder scores of its hits in output order.  Re-emits input_ids / boxes / mask (and the hit_*
        arrays) in the reranked order on the device; returns (PackedInputs, order, kept[, visual input])."""
        order, kept, _ = rerank_order(scores, plan.t["topk_cnt"], self.rerank_filter_tresh, self.rerank_max_chunk_num,
                                      self.rerank_min_chunk_num)
        plan.set_emit_order(order, kept)
        plan.launch()
        visual = pages.pack(plan.t["hit_i"][1], plan.t["hit_rect"], kept, **pack_options) if pages is not None else None
        packed = plan.finish()
        return (packed, order, kept) if pages is None else (packed, order, kept, visual)
