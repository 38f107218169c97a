"""Tensor-level entry points of the B200 retrieval path (thin host code over librdv.so).

Every function here launches hand-written sm_100a kernels through the C ABI in include/rdv.h on
torch's current CUDA stream.  PyTorch is used for device memory and streams only.  Inputs must live on
a CUDA device: there is no CPU path (use oracle/ for a CPU checker).
"""
from __future__ import annotations

import collections
import ctypes
from typing import List, NamedTuple, Sequence

import numpy as np
import torch

from . import _lib

_lib_fn = _lib.lib


def _stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def _require_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise RuntimeError("%s must be a CUDA tensor: rag_docvqa_b200 has no CPU fallback "
                           "(got device %s)" % (what, t.device))


def _f32_contig_aligned(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.float32:
        t = t.float()
    if not t.is_contiguous():
        t = t.contiguous()
    if t.data_ptr() % 16:
        t = t.clone(memory_format=torch.contiguous_format)
    return t


class _Workspace:
    """Per (device, stream) zero-initialised int32 scratch that the kernels leave zeroed."""
    _cache = {}

    @classmethod
    def zeros_i32(cls, device, n: int) -> torch.Tensor:
        key = (device.index, _stream_ptr(device))
        buf = cls._cache.get(key)
        if buf is None or buf.numel() < n:
            buf = torch.zeros(max(n, 1024), dtype=torch.int32, device=device)
            cls._cache[key] = buf
        return buf


class EmbeddingCache:
    """Device copies of HOST document embeddings, kept between calls (least recently used first out): MP-DocVQA asks many
    questions about one document, and a read-once streaming op cannot win across PCIe -- the second question about a
    document should not cross it again.  A document is recognised by its tensor's storage address, shape, dtype and version
    counter (in-place torch writes bump it; writes through numpy views do not -- callers that do that must not use the
    cache)."""

    def __init__(self, budget_bytes: int, device):
        from collections import OrderedDict
        self.budget, self.device = int(budget_bytes), torch.device(device)
        self.entries = OrderedDict()
        self.bytes = 0
        self.hits = self.misses = 0

    def resident(self, host_docs: Sequence[torch.Tensor]) -> List[torch.Tensor]:
        out = []
        for e in host_docs:
            if e.is_cuda or e.shape[0] == 0:
                out.append(e if e.is_cuda else e.to(self.device))
                continue
            key = (e.data_ptr(), tuple(e.shape), e.dtype, e._version)
            hit = self.entries.get(key)
            if hit is not None:
                self.entries.move_to_end(key)
                self.hits += 1
                out.append(hit)
                continue
            self.misses += 1
            t = e.to(self.device, non_blocking=True)
            nbytes = t.numel() * t.element_size()
            if nbytes <= self.budget:
                self.entries[key] = t
                self.bytes += nbytes
                while self.bytes > self.budget:
                    _, old = self.entries.popitem(last=False)
                    self.bytes -= old.numel() * old.element_size()
            out.append(t)
        return out


class ScoreTopK(NamedTuple):
    similarities: List[torch.Tensor]   # List[B] of (n_b,) fp32 views into `sims`
    sims: torch.Tensor                 # (N,) fp32, all documents back to back
    topk_idx: torch.Tensor             # (B, k) int32, -1 padded, rank order
    topk_val: torch.Tensor             # (B, k) fp32, -inf padded
    topk_cnt: torch.Tensor             # (B,) int32 = min(k, n_b)
    sizes: List[int]


TILE_DTYPE = np.dtype([("src", "<u8"), ("sims_off", "<i8"), ("rows", "<i4"), ("doc", "<i4"),
                       ("doc_rows", "<i4"), ("reserved", "<i4")])      # rdv_tile_desc, 32 bytes
CTA_DTYPE = np.dtype([("src", "<u8"), ("sims_off", "<i4"), ("doc_rows", "<i4"), ("doc", "<i4"), ("part", "<i2"),
                      ("nparts", "<i2"), ("r0", "<i4"), ("r1", "<i4")])  # rdv_cta_desc, 32 bytes


class DocTable(NamedTuple):
    """Device-side description of a ragged batch of documents (see rdv_score_topk_f32 / rdv_score_topk_cluster_f32)."""
    desc: torch.Tensor     # uint8 blob on device: row_off[B+1] i64 | pad | tiles[T] rdv_tile_desc | ctas[n_ctas] rdv_cta_desc
    keepalive: tuple       # tensors whose storage the descriptors point into
    B: int
    d: int
    sizes: List[int]
    total_rows: int
    total_tiles: int
    tile_rows: int
    max_rows: int
    algo: int
    tiles_offset: int
    ctas_offset: int
    n_ctas: int            # descriptors of the cluster kernels (0: the batch is outside their limits)
    cluster: int           # ... and the cluster size they were packed for

    def pointers(self):
        """(d_tiles, d_row_off)"""
        base = self.desc.data_ptr()
        return base + self.tiles_offset, base

    def cluster_pointers(self):
        """(d_ctas, n_ctas, cluster) for the cluster kernels"""
        return self.desc.data_ptr() + self.ctas_offset, self.n_ctas, self.cluster

    def cluster_fits(self, k: int) -> bool:
        """Do the one-launch cluster kernels apply to this batch (their limits; a cluster table was built)?"""
        return bool(self.B and self.n_ctas and self.algo != _lib.SCORE_TMA and 1 <= int(k) <= int(_lib_fn.rdv_cluster_max_k()))

    def use_cluster(self, k: int) -> bool:
        """rdv_retrieve_plan: ... and are they the better choice?  (Measured in round 2: no -- see csrc/retrieve_cluster.cu.)"""
        if not self.cluster_fits(k):
            return False
        out = ctypes.c_int32()
        _lib.check(_lib_fn.rdv_retrieve_plan(self.total_rows, self.max_rows, self.B, self.d, int(k), ctypes.byref(out)))
        return bool(out.value)


def plan_score(total_rows: int, d: int, algo: int = _lib.SCORE_AUTO):
    algo_out, tile_rows = ctypes.c_int32(), ctypes.c_int32()
    _lib.check(_lib_fn.rdv_score_plan(total_rows, d, algo, ctypes.byref(algo_out), ctypes.byref(tile_rows)))
    return algo_out.value, tile_rows.value


_TABLE_CACHE = collections.OrderedDict()


def _table_from_pointers(ptrs: np.ndarray, sizes: np.ndarray, keep, d: int, device, tile_rows: int, algo: int,
                         plan_k: int = 8) -> DocTable:
    """rdv_build_doc_table into a pinned blob (row_off | pad | tiles | cluster descriptors) + ONE H2D copy.
    The blob is a pure function of (pointers, sizes, d, tile_rows, algo), so the device copies of the last few are kept
    (per device and stream): a caller that comes back with the same resident documents -- more questions about one
    document, a serving loop over a fixed index -- pays neither the cutting nor the upload again."""
    key = (ptrs.tobytes(), sizes.tobytes(), d, tile_rows, algo, plan_k, torch.device(device).index, _stream_ptr(device))
    hit = _TABLE_CACHE.get(key)
    if hit is not None:
        _TABLE_CACHE.move_to_end(key)
        return hit._replace(keepalive=keep)
    table = _table_from_pointers_uncached(ptrs, sizes, keep, d, device, tile_rows, algo, plan_k)
    _TABLE_CACHE[key] = table._replace(keepalive=())          # the cache must not keep the documents alive
    while len(_TABLE_CACHE) > 32:
        _TABLE_CACHE.popitem(last=False)
    return table


def _table_from_pointers_uncached(ptrs, sizes, keep, d, device, tile_rows, algo, plan_k) -> DocTable:
    B = len(sizes)
    total_rows = int(sizes.sum()) if B else 0
    algo, planned_rows = plan_score(total_rows, d, algo)
    if tile_rows <= 0:
        tile_rows = planned_rows
    p_sizes = sizes.ctypes.data
    T = int(_lib_fn.rdv_count_tiles(p_sizes, B, tile_rows)) if B else 0
    if T < 0 or T >= 2 ** 31 - 1:
        raise ValueError("bad document sizes / too many tiles for one launch")
    tiles_offset = (8 * (B + 1) + 31) // 32 * 32
    ctas_offset = tiles_offset + 32 * max(T, 1)
    # the cluster kernels' view (tiles packed into clusters, a document never straddling one) where their limits allow
    n_ctas, cluster, slice_rows = 0, 0, 0
    if B and total_rows < 2 ** 31 and algo != _lib.SCORE_TMA and int(sizes.max()) <= int(_lib_fn.rdv_cluster_max_rows(16)):
        c_cluster, c_slice, c_n = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int64()
        with torch.cuda.device(device):      # asked for the variant that also gathers with k <= 8 (the common one, and the
            # one with the least room: a plan that fits it fits the score + top-k variant as well)
            _lib.check(_lib_fn.rdv_cluster_plan(p_sizes, B, d, plan_k, 1, ctypes.byref(c_cluster), ctypes.byref(c_slice),
                                                ctypes.byref(c_n)))
        cluster, slice_rows, n_ctas = c_cluster.value, c_slice.value, int(c_n.value)
    host = torch.empty(ctas_offset + 32 * max(n_ctas, 1), dtype=torch.uint8, pin_memory=True)
    base = host.data_ptr()
    max_rows = ctypes.c_int32(0)
    if B:
        _lib.check(_lib_fn.rdv_build_doc_table(ptrs.ctypes.data, p_sizes, B, d, tile_rows, base, base + tiles_offset, T,
                                               ctypes.byref(max_rows)))
        if n_ctas:
            _lib.check(_lib_fn.rdv_build_cluster_table(ptrs.ctypes.data, p_sizes, B, d, cluster, slice_rows, base + ctas_offset,
                                                       n_ctas))
    else:
        host[:8] = 0
    desc = host.to(device, non_blocking=True)
    return DocTable(desc, keep, B, d, sizes.tolist(), total_rows, T, tile_rows, int(max_rows.value), algo, tiles_offset,
                    ctas_offset, n_ctas, cluster)


def build_doc_table(docs: Sequence[torch.Tensor], d: int, device, tile_rows: int = 0,
                    algo: int = _lib.SCORE_AUTO) -> DocTable:
    """Cuts a ragged batch of DEVICE tensors into row tiles and uploads offsets + tile descriptors in ONE pinned
    H2D copy (the cutting itself is rdv_build_doc_table, in C)."""
    B = len(docs)
    keep = []
    sizes = np.empty(B, dtype=np.int64)
    ptrs = np.empty(B, dtype=np.uint64)
    for b, t in enumerate(docs):
        if t.dim() != 2 or (t.shape[0] and t.shape[1] != d):
            raise ValueError("document %d: expected (n, %d) embeddings, got %s" % (b, d, tuple(t.shape)))
        if t.shape[0]:
            _require_cuda(t, "text_embeddings[%d]" % b)
            t = _f32_contig_aligned(t)
        keep.append(t)
        sizes[b] = t.shape[0]
        ptrs[b] = t.data_ptr() if t.shape[0] else 0
    return _table_from_pointers(ptrs, sizes, tuple(keep), d, device, tile_rows, algo)


def upload_doc_table(host_docs: Sequence[torch.Tensor], d: int, device, tile_rows: int = 0,
                     algo: int = _lib.SCORE_AUTO) -> DocTable:
    """HOST documents (the bench's e2e case; in the reference they already live on the GPU).

    Pinned host tensors are NOT copied: page-locked memory is mapped into the device's address space (UVA), so the
    tile descriptors point straight at the host rows and the score kernel streams them over PCIe itself -- each row
    is read exactly once, so a staging copy would only add the copy engine's per-transfer latency (measured on B200:
    64 per-document copies of a 32 MB C2 batch 1.47 ms, one 32 MB copy 0.63 ms).  Pageable tensors go through one
    packed device buffer, one cudaMemcpyAsync per document issued from C (rdv_upload_docs_f32)."""
    B = len(host_docs)
    if B and all(t.is_pinned() and t.dtype == torch.float32 and t.is_contiguous() and t.data_ptr() % 16 == 0
                 for t in host_docs if t.shape[0]):
        sizes = np.fromiter((t.shape[0] for t in host_docs), dtype=np.int64, count=B)
        for b, t in enumerate(host_docs):
            if t.dim() != 2 or (t.shape[0] and t.shape[1] != d):
                raise ValueError("document %d: expected (n, %d) embeddings, got %s" % (b, d, tuple(t.shape)))
        ptrs = np.fromiter((t.data_ptr() if t.shape[0] else 0 for t in host_docs), dtype=np.uint64, count=B)
        return _table_from_pointers(ptrs, sizes, tuple(host_docs), d, device, tile_rows, algo)
    docs = []
    sizes = np.empty(B, dtype=np.int64)
    hptrs = np.empty(B, dtype=np.uint64)
    for b, t in enumerate(host_docs):
        if t.dim() != 2 or (t.shape[0] and t.shape[1] != d):
            raise ValueError("document %d: expected (n, %d) embeddings, got %s" % (b, d, tuple(t.shape)))
        if t.is_cuda:
            raise ValueError("upload_doc_table takes host tensors")
        if t.dtype != torch.float32 or not t.is_contiguous():
            t = t.float().contiguous()
        docs.append(t)
        sizes[b] = t.shape[0]
        hptrs[b] = t.data_ptr() if t.shape[0] else 0
    total = int(sizes.sum()) if B else 0
    packed = torch.empty((max(total, 1), d), dtype=torch.float32, device=device)
    if B:
        _lib.check(_lib_fn.rdv_upload_docs_f32(hptrs.ctypes.data, sizes.ctypes.data, B, d, packed.data_ptr(),
                                               _stream_ptr(device)))
    row0 = np.zeros(B, dtype=np.int64)
    if B > 1:
        np.cumsum(sizes[:-1], out=row0[1:])
    dptrs = (np.uint64(packed.data_ptr()) + (row0 * (d * 4)).astype(np.uint64)) * (sizes > 0).astype(np.uint64)
    return _table_from_pointers(dptrs, sizes, (packed, tuple(docs)), d, device, tile_rows, algo)


def score_topk_table(table: DocTable, questions: torch.Tensor, k: int, cluster=None) -> ScoreTopK:
    """Score + per-document top-k on an uploaded DocTable: ONE launch (a thread-block cluster per document,
    rdv_score_topk_cluster_f32) for batches of short documents, else the streaming kernel + the selection kernel
    (rdv_score_topk_f32).  `cluster`: None = rdv_retrieve_plan decides, True / False = forced."""
    device = questions.device
    B, d = table.B, table.d
    q = _f32_contig_aligned(questions)
    sims_base = torch.empty(max(table.total_rows, 1), dtype=torch.float32, device=device)    # never a null pointer
    sims = sims_base[:table.total_rows]
    topk_idx = torch.empty((B, k), dtype=torch.int32, device=device)
    topk_val = torch.empty((B, k), dtype=torch.float32, device=device)
    topk_cnt = torch.empty((B,), dtype=torch.int32, device=device)
    if B and (table.use_cluster(k) if cluster is None else cluster):
        d_ctas, n_ctas, cluster_size = table.cluster_pointers()
        _lib.check(_lib_fn.rdv_score_topk_cluster_f32(
            d_ctas, n_ctas, cluster_size, q.data_ptr(), B, d, k, table.max_rows, sims_base.data_ptr(), topk_idx.data_ptr(),
            topk_val.data_ptr(), topk_cnt.data_ptr(), _stream_ptr(device)))
    elif B:
        p_tiles, p_row = table.pointers()
        _lib.check(_lib_fn.rdv_score_topk_f32(
            p_tiles, table.total_tiles, table.tile_rows, table.algo, p_row, q.data_ptr(), B, d, k,
            table.max_rows, sims_base.data_ptr(), topk_idx.data_ptr(), topk_val.data_ptr(),
            topk_cnt.data_ptr(), _stream_ptr(device)))
    views = list(torch.split(sims, table.sizes)) if B else []
    return ScoreTopK(views, sims, topk_idx, topk_val, topk_cnt, table.sizes)


def score_topk(text_embeddings: Sequence[torch.Tensor], question_embeddings: torch.Tensor, k: int,
               tile_rows: int = 0, algo: int = _lib.SCORE_AUTO, cluster=None) -> ScoreTopK:
    """Cosine score of question b against every chunk of document b + per-document top-k.

    Replaces Retriever._get_similarities + torch.topk (reference src/_modules.py:1978-1997, 2015-2016).
    text_embeddings: List[B] of (n_b, d) CUDA fp32 tensors (n_b may be 0); question_embeddings: (B, d).
    """
    _require_cuda(question_embeddings, "question_embeddings")
    if question_embeddings.dim() != 2 or question_embeddings.shape[0] != len(text_embeddings):
        raise ValueError("question_embeddings must be (B, d) with B == len(text_embeddings)")
    if k < 1:
        raise ValueError("k must be >= 1")
    d = question_embeddings.shape[1]
    with torch.cuda.device(question_embeddings.device):
        table = build_doc_table(text_embeddings, d, question_embeddings.device, tile_rows, algo)
        return score_topk_table(table, question_embeddings, k, cluster=cluster)


# ------------------------------------------------------------------------------------------------
# mean pooling (a1)
# ------------------------------------------------------------------------------------------------
def mean_pooling(embs: torch.Tensor, attention_mask: torch.Tensor, normalise: bool = False,
                 out_bf16: bool = False, return_norm: bool = False):
    """Drop-in for mean_pooling (reference src/_model_utils.py:49-61): masked mean over tokens.

    embs (n, L, d) CUDA fp32, attention_mask (n, L) integer/bool/float -> (n, d) fp32.
    Extras (not in the reference signature, defaults preserve behaviour): `normalise` fuses the L2
    normalisation of the pooled row; `out_bf16` returns a bf16 copy as well (corpus shards);
    `return_norm` returns the pooled rows' L2 norms.
    """
    _require_cuda(embs, "embs")
    if embs.dim() != 3 or attention_mask.shape != embs.shape[:2]:
        raise ValueError("mean_pooling: embs must be (n, L, d) and attention_mask (n, L)")
    device = embs.device
    n, L, d = embs.shape
    e = _f32_contig_aligned(embs)
    m = attention_mask.to(device=device, dtype=torch.int64).contiguous()
    out = torch.empty((n, d), dtype=torch.float32, device=device)
    out16 = torch.empty((n, d), dtype=torch.bfloat16, device=device) if out_bf16 else None
    norm = torch.empty((n,), dtype=torch.float32, device=device) if return_norm else None
    with torch.cuda.device(device):
        _lib.check(_lib_fn.rdv_mean_pool_f32(
            e.data_ptr(), m.data_ptr(), n, L, d, 1 if normalise else 0, out.data_ptr(),
            out16.data_ptr() if out16 is not None else None, norm.data_ptr() if norm is not None else None,
            _stream_ptr(device)))
    extras = tuple(x for x in (out16, norm) if x is not None)
    return (out, *extras) if extras else out


# ------------------------------------------------------------------------------------------------
# pooled-patch visual retrieval (north_star's wording of BASELINE.json configs[3])
# ------------------------------------------------------------------------------------------------
_POOLED_OFFSETS = {}


def _pooled_offsets(n_strips: tuple, L: int, kk: int, device):
    """Segment offsets of pooled_patch_topk's three selections for a batch shape (kept: they only depend on the shape)."""
    key = (n_strips, L, kk, device.index)
    hit = _POOLED_OFFSETS.get(key)
    if hit is None:
        if len(_POOLED_OFFSETS) > 64:
            _POOLED_OFFSETS.clear()
        doc_off = np.zeros(len(n_strips) + 1, dtype=np.int64)
        np.cumsum(n_strips, out=doc_off[1:])
        total = int(doc_off[-1])
        hit = (torch.arange(0, (total + 1) * L, L, dtype=torch.int64, device=device),
               torch.from_numpy(doc_off * kk).to(device), torch.from_numpy(doc_off).to(device))
        _POOLED_OFFSETS[key] = hit
    return hit


class PooledPatchTopK(NamedTuple):
    similarities: List[torch.Tensor]   # List[B] of (n_b * L,) fp32: cosine of every patch vector, strip-major
    patch_idx: torch.Tensor            # (B, k) int32 flat patch index (strip * L + patch), -1 padded, rank order
    patch_val: torch.Tensor            # (B, k) fp32
    patch_cnt: torch.Tensor            # (B,) int32
    strip_scores: List[torch.Tensor]   # List[B] of (n_b,) fp32: best patch of every strip
    strip_idx: torch.Tensor            # (B, k_strips) int32 strip index, -1 padded, rank order
    strip_val: torch.Tensor
    strip_cnt: torch.Tensor
    question: torch.Tensor             # (B, d) fp32 pooled question vectors


def pooled_patch_topk(patch_embeddings: Sequence[torch.Tensor], question_embeddings: torch.Tensor, k: int,
                      question_mask: torch.Tensor = None, k_strips: int = None) -> PooledPatchTopK:
    """Pooled-patch retrieval over page strips: the question's encoder tokens are mean-pooled (mean_pooling,
    src/_model_utils.py:49-61; all tokens when no mask is given), EVERY patch vector of every strip -- the (n_b, L, d)
    encoder outputs ImageEncoder returns (src/_modules.py:1627-1666), 102 400 vectors for 50 strips -- is scored against it with
    Retriever._get_similarities' cosine (src/_modules.py:1990-1993: eps on the product of the norms), the k best patches per
    document are selected, a strip is scored by its best patch and the k_strips best strips are selected (torch.topk,
    src/_modules.py:2408; lowest index first on ties).  Four launches: pooling, the streaming score kernel over all documents,
    the top-k of every strip, and per document the top-k of its strips' candidates + the strip scores + their top-k
    (rdv_pooled_select_f32); every patch vector is read from HBM exactly once."""
    _require_cuda(question_embeddings, "question_embeddings")
    device = question_embeddings.device
    B = len(patch_embeddings)
    if question_embeddings.dim() != 3 or question_embeddings.shape[0] != B:
        raise ValueError("pooled_patch_topk: question_embeddings must be (B, Lq, d) with B == len(patch_embeddings)")
    d = question_embeddings.shape[2]
    if question_mask is None:
        question_mask = torch.ones(question_embeddings.shape[:2], dtype=torch.int64, device=device)
    k_strips = int(k if k_strips is None else k_strips)
    with torch.cuda.device(device):
        q = mean_pooling(question_embeddings, question_mask)                       # (B, d)
        L = None
        flat, n_strips = [], []
        for b, p in enumerate(patch_embeddings):
            if p.dim() != 3 or p.shape[2] != d:
                raise ValueError("pooled_patch_topk: document %d: expected (n, L, %d) patch embeddings, got %s" % (b, d, tuple(p.shape)))
            if p.shape[0]:
                _require_cuda(p, "patch_embeddings[%d]" % b)
                if L is None:
                    L = int(p.shape[1])
                elif int(p.shape[1]) != L:
                    raise ValueError("pooled_patch_topk: strips of %d and %d patches in one batch" % (L, int(p.shape[1])))
            n_strips.append(int(p.shape[0]))
            flat.append(_f32_contig_aligned(p).reshape(-1, d))                     # a view: nothing is copied
        L = L or 1
        k = int(k)
        table = build_doc_table(flat, d, device)
        sims = score_table(table, q)                                               # every patch vector, read once
        total_strips = sum(n_strips)
        sizes = [n * L for n in n_strips]
        similarities = list(torch.split(sims, sizes)) if B else []
        fill = total_strips == 0                                                    # no strip at all: nothing launches
        patch_idx = (torch.full if fill else torch.empty)((B, k), *((-1,) if fill else ()), dtype=torch.int32, device=device)
        patch_val = (torch.full if fill else torch.empty)((B, k), *((float("-inf"),) if fill else ()), dtype=torch.float32, device=device)
        patch_cnt = torch.zeros((B,), dtype=torch.int32, device=device) if fill else torch.empty((B,), dtype=torch.int32, device=device)
        strip_scores = [sims.new_empty(0) for _ in range(B)]
        s_idx = (torch.full if fill else torch.empty)((B, k_strips), *((-1,) if fill else ()), dtype=torch.int32, device=device)
        s_val = (torch.full if fill else torch.empty)((B, k_strips), *((float("-inf"),) if fill else ()), dtype=torch.float32, device=device)
        s_cnt = torch.zeros((B,), dtype=torch.int32, device=device) if fill else torch.empty((B,), dtype=torch.int32, device=device)
        if total_strips:
            # two-level selection: a document's 102 400 scores are 50 segments of 2048, so the k best of every STRIP come
            # first (one block per strip, register-resident: 400 blocks instead of 8), then per document the k best of its
            # strips' candidates; a strip's own score is its rank-0 candidate (torch.max: NaN greatest, as torch.topk ranks
            # it).  Both levels, the index arithmetic and the strips' top-k are two launches of ONE C call.
            kk = min(k, L)
            strip_off, _, strip_off_doc = _pooled_offsets(tuple(n_strips), L, kk, device)
            ws_idx = torch.empty((total_strips, kk), dtype=torch.int32, device=device)
            ws_val = torch.empty((total_strips, kk), dtype=torch.float32, device=device)
            ws_cnt = torch.empty((total_strips,), dtype=torch.int32, device=device)
            best = torch.empty((total_strips,), dtype=torch.float32, device=device)
            _lib.check(_lib_fn.rdv_pooled_select_f32(
                sims.data_ptr(), strip_off.data_ptr(), total_strips, L, strip_off_doc.data_ptr(), B, max(n_strips), k, k_strips,
                ws_idx.data_ptr(), ws_val.data_ptr(), ws_cnt.data_ptr(), patch_idx.data_ptr(), patch_val.data_ptr(),
                patch_cnt.data_ptr(), best.data_ptr(), s_idx.data_ptr(), s_val.data_ptr(), s_cnt.data_ptr(), _stream_ptr(device)))
            strip_scores = list(torch.split(best, n_strips))
    return PooledPatchTopK(similarities, patch_idx, patch_val, patch_cnt, strip_scores, s_idx, s_val, s_cnt, q)



# ------------------------------------------------------------------------------------------------
# MaxSim late interaction (a5)
# ------------------------------------------------------------------------------------------------
def split_tf32(x: torch.Tensor, normalise: bool = False):
    """fp32 (..., d) rows -> (hi, lo) with hi = tf32(x), lo = x - hi exactly (optionally after F.normalize)."""
    _require_cuda(x, "x")
    d = x.shape[-1]
    xf = _f32_contig_aligned(x)
    hi = torch.empty_like(xf)
    lo = torch.empty_like(xf)
    with torch.cuda.device(x.device):
        _lib.check(_lib_fn.rdv_rows_split_tf32(xf.data_ptr(), xf.numel() // d, d, 1 if normalise else 0, hi.data_ptr(),
                                               lo.data_ptr(), _stream_ptr(x.device)))
    return hi, lo


def late_interaction_tf32x3(query: torch.Tensor, patches: torch.Tensor) -> torch.Tensor:
    """late_interaction at fp32 grade on the tensor pipe: operands split into tf32 hi + lo parts, three tcgen05
    kind::tf32 products accumulated in fp32 in TMEM (csrc/tc_tf32.cu).  Same contract as late_interaction."""
    _require_cuda(query, "query")
    _require_cuda(patches, "patches")
    if query.dim() == 2:
        query = query.unsqueeze(0)
    n, Lp, d = patches.shape
    Lq = query.shape[1]
    device = patches.device
    out = torch.empty((n,), dtype=torch.float32, device=device)
    if n == 0:
        return out
    q_hi, q_lo = split_tf32(query[0], normalise=True)
    p_hi, p_lo = split_tf32(patches, normalise=True)
    tiles = (Lq + int(_lib_fn.rdv_tc_tile_m()) - 1) // int(_lib_fn.rdv_tc_tile_m())
    partial = torch.empty((n * tiles,), dtype=torch.float32, device=device)
    with torch.cuda.device(device):
        _lib.check(_lib_fn.rdv_maxsim_tf32x3_tc(q_hi.data_ptr(), q_lo.data_ptr(), p_hi.data_ptr(), p_lo.data_ptr(), n, Lq,
                                                Lp, d, partial.data_ptr(), out.data_ptr(), _stream_ptr(device)))
    return out


def late_interaction(query: torch.Tensor, patches: torch.Tensor, mode: str = "auto") -> torch.Tensor:
    """Drop-in for late_interaction (reference src/utils.py:442-458), fp32 parity mode.

    query (1, Lq, d), patches (n, Lp, d) CUDA fp32 -> (n,) scores = sum_i max_j cos(q_i, p_nj).
    mode: "ffma" = CUDA-core fp32 kernel (csrc/maxsim.cu), ~1e-7 from float64; "tf32x3" = split tf32 products on
    the tensor pipe (csrc/tc_tf32.cu), 5.4x faster at C4, scores ~d * 2.1e-9 relative low (1.6e-6 at d = 768, the
    tensor core's accumulator rounding); "auto" = tf32x3 when the contraction is large enough to fill the tensor
    pipe and d <= 2048 keeps that error well inside the 1e-5 parity bar, else ffma.
    """
    _require_cuda(query, "query")
    _require_cuda(patches, "patches")
    if query.dim() == 2:
        query = query.unsqueeze(0)
    if query.dim() != 3 or query.shape[0] != 1 or patches.dim() != 3 or patches.shape[2] != query.shape[2]:
        raise ValueError("late_interaction: query must be (1, Lq, d) and patches (n, Lp, d)")
    if mode not in ("auto", "ffma", "tf32x3"):
        raise ValueError("late_interaction: unknown mode %r" % (mode,))
    device = patches.device
    n, Lp, d = patches.shape
    Lq = query.shape[1]
    if mode == "tf32x3" or (mode == "auto" and n * Lq * Lp * d >= (1 << 28) and Lq >= 64 and Lp >= 64 and d <= 2048):
        return late_interaction_tf32x3(query, patches)
    q = _f32_contig_aligned(query[0])
    p = _f32_contig_aligned(patches)
    out = torch.empty((n,), dtype=torch.float32, device=device)
    if n == 0:
        return out
    with torch.cuda.device(device):
        s = _stream_ptr(device)
        inv_q = torch.empty((Lq,), dtype=torch.float32, device=device)
        inv_p = torch.empty((n * Lp,), dtype=torch.float32, device=device)
        _lib.check(_lib_fn.rdv_row_inv_norm_f32(q.data_ptr(), Lq, d, inv_q.data_ptr(), s))
        _lib.check(_lib_fn.rdv_row_inv_norm_f32(p.data_ptr(), n * Lp, d, inv_p.data_ptr(), s))
        tiles_i = int(_lib_fn.rdv_maxsim_tiles_i(Lq))
        partial = torch.empty((n * tiles_i,), dtype=torch.float32, device=device)
        counter = _Workspace.zeros_i32(device, n)
        # grid.y is limited to 65535 strips per launch
        for lo in range(0, n, 65535):
            hi = min(n, lo + 65535)
            _lib.check(_lib_fn.rdv_maxsim_f32(
                q.data_ptr(), p.data_ptr() + lo * Lp * d * 4, inv_q.data_ptr(), inv_p.data_ptr() + lo * Lp * 4,
                hi - lo, Lq, Lp, d, partial.data_ptr() + lo * tiles_i * 4, counter.data_ptr(),
                out.data_ptr() + lo * 4, s))
    return out


# ------------------------------------------------------------------------------------------------
# shard merge (corpus mode)
# ------------------------------------------------------------------------------------------------
def topk_merge(cand_val: torch.Tensor, cand_idx: torch.Tensor, k: int):
    """(Q, m) candidates (global ids, <0 = empty) -> (Q, k) values / int64 ids by (score desc, id asc)."""
    _require_cuda(cand_val, "cand_val")
    device = cand_val.device
    Q, m = cand_val.shape
    v = cand_val.float().contiguous()
    i = cand_idx.to(torch.int64).contiguous()
    out_v = torch.empty((Q, k), dtype=torch.float32, device=device)
    out_i = torch.empty((Q, k), dtype=torch.int64, device=device)
    with torch.cuda.device(device):
        _lib.check(_lib_fn.rdv_topk_merge(v.data_ptr(), i.data_ptr(), Q, m, k, out_v.data_ptr(), out_i.data_ptr(),
                                          _stream_ptr(device)))
    return out_v, out_i


def topk_segments(scores: Sequence[torch.Tensor], k: int):
    """Per-document top-k (score desc, lowest index first, NaN greatest) of existing score vectors.
    Replaces torch.topk at reference src/_modules.py:2408.  Returns (idx (B,k) int32, val, cnt)."""
    B = len(scores)
    if B == 0:
        raise ValueError("topk_segments: empty batch")
    device = scores[0].device
    _require_cuda(scores[0], "scores")
    sizes = [int(s.shape[0]) for s in scores]
    flat = torch.cat([s.float().reshape(-1) for s in scores]) if sum(sizes) else torch.empty(0, device=device)
    row_off = np.zeros(B + 1, dtype=np.int64)
    np.cumsum(sizes, out=row_off[1:])
    off_d = torch.from_numpy(row_off).pin_memory().to(device, non_blocking=True)
    idx = torch.empty((B, k), dtype=torch.int32, device=device)
    val = torch.empty((B, k), dtype=torch.float32, device=device)
    cnt = torch.empty((B,), dtype=torch.int32, device=device)
    with torch.cuda.device(device):
        _lib.check(_lib_fn.rdv_topk_segments_f32(flat.data_ptr(), off_d.data_ptr(), B, k, max(sizes), idx.data_ptr(),
                                                 val.data_ptr(), cnt.data_ptr(), _stream_ptr(device)))
    return idx, val, cnt


# ------------------------------------------------------------------------------------------------
# bf16 tensor-core modes (tcgen05 / TMEM / TMA)
# ------------------------------------------------------------------------------------------------
def rows_to_bf16(x: torch.Tensor, normalise: bool = False, return_inv_norm: bool = False):
    """fp32 (..., d) rows -> bf16 copy (optionally L2-normalised first, F.normalize semantics)."""
    _require_cuda(x, "x")
    d = x.shape[-1]
    xf = _f32_contig_aligned(x)
    rows = xf.numel() // d
    out = torch.empty(xf.shape, dtype=torch.bfloat16, device=x.device)
    inv = torch.empty((rows,), dtype=torch.float32, device=x.device) if return_inv_norm else None
    with torch.cuda.device(x.device):
        _lib.check(_lib_fn.rdv_rows_to_bf16(xf.data_ptr(), rows, d, 1 if normalise else 0, out.data_ptr(),
                                            inv.data_ptr() if inv is not None else None, _stream_ptr(x.device)))
    return (out, inv) if return_inv_norm else out


def bf16_inv_norm(x: torch.Tensor) -> torch.Tensor:
    _require_cuda(x, "x")
    if x.dtype != torch.bfloat16 or not x.is_contiguous():
        raise ValueError("bf16_inv_norm: expected a contiguous bf16 matrix")
    rows, d = x.shape
    inv = torch.empty((rows,), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib_fn.rdv_bf16_inv_norm(x.data_ptr(), rows, d, inv.data_ptr(), _stream_ptr(x.device)))
    return inv


def late_interaction_bf16(query: torch.Tensor, patches: torch.Tensor) -> torch.Tensor:
    """MaxSim fast mode: same contract as late_interaction, bf16 operands on the tcgen05 tensor cores
    (fp32 accumulation in TMEM).  Reported against the fp32 mode as a score error / top-k recall."""
    _require_cuda(query, "query")
    _require_cuda(patches, "patches")
    if query.dim() == 2:
        query = query.unsqueeze(0)
    n, Lp, d = patches.shape
    Lq = query.shape[1]
    device = patches.device
    out = torch.empty((n,), dtype=torch.float32, device=device)
    if n == 0:
        return out
    qn = rows_to_bf16(query[0], normalise=True)
    pn = rows_to_bf16(patches, normalise=True)
    tiles = (Lq + int(_lib_fn.rdv_tc_tile_m()) - 1) // int(_lib_fn.rdv_tc_tile_m())
    partial = torch.empty((n * tiles,), dtype=torch.float32, device=device)
    with torch.cuda.device(device):
        _lib.check(_lib_fn.rdv_maxsim_bf16_tc(qn.data_ptr(), pn.data_ptr(), n, Lq, Lp, d, partial.data_ptr(),
                                              out.data_ptr(), _stream_ptr(device)))
    return out


def score_table(table: DocTable, questions: torch.Tensor, out: torch.Tensor = None) -> torch.Tensor:
    """Scores only (rdv_score_f32): every similarity of the batch, (N,) fp32.  The selection then runs in
    topk_segments or inside the gather kernel (DocStore.prepare_gather(..., sims=...))."""
    device = questions.device
    q = _f32_contig_aligned(questions)
    sims = out if out is not None else torch.empty(table.total_rows, dtype=torch.float32, device=device)
    if table.B and table.total_tiles:
        p_tiles, _ = table.pointers()
        _lib.check(_lib_fn.rdv_score_f32(p_tiles, table.total_tiles, table.tile_rows, table.algo, q.data_ptr(), table.B,
                                         table.d, sims.data_ptr(), _stream_ptr(device)))
    return sims
