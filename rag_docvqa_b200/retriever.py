"""Drop-in replacements for the reference's retrieval classes, backed by the sm_100a kernels.

    Retriever(config).retrieve(...)         reference src/_modules.py:1967-2180
    VisualRetriever(config).retrieve(...)   reference src/_modules.py:2183-2464

Same constructor (one flat config dict, same keys), same call signatures, same output format
(nested Python lists, PIL crops, per-document similarity tensors), so src/RAGVT5.py:105,244-252 and
src/RAGPix2Struct.py:83,157-164 run unchanged (see rag_docvqa_b200/compat and INTEGRATION.md).

What moved to the GPU: every similarity (fused cosine / MaxSim), the per-document top-k (one launch for
the whole batch instead of >= 6 launches + k host syncs per document), and -- through
`retrieve_packed` -- the gather into the generator's input tensors.  What stays on the host: building the
Python list/PIL view of the <= k hits per document, which only touches the hits (the reference walks
every word of every chunk, src/_modules.py:2032-2050).

Differences from the reference, all within north_star's contract:
  * ties are broken by LOWEST index (torch.topk's tie order is unspecified);
  * VisualRetriever returns crops / page ids in sorted (group, rectangle) order (the reference's order is
    Python-set iteration order, src/_modules.py:2428,2445).
"""
from __future__ import annotations

import ctypes
import gc
from typing import Any, Dict, List, Optional, Sequence, Union

import numpy as np
import torch

from . import functional as F

try:                                   # C helper for the nested-list view (csrc/hostlists.c); same semantics as
    from . import _hostlists           # the Python walk below, which stays as the general path and the spec
except ImportError:                    # pragma: no cover - built by rag_docvqa_b200.build
    _hostlists = None

_LAYOUT_MAP_4 = {0: "title", 1: "text", 2: "figure", 3: "table"}   # src/_modules.py:308-313, 636-641


def get_layout_model_map(config: dict) -> dict:
    """reference src/_modules.py:246-253."""
    return dict(_LAYOUT_MAP_4) if config.get("layout_model") in ("YOLO", "DIT") else {1: "text"}


class StatComponent:
    """Counter plumbing every reference component carries (src/_modules.py:178-243); `stats` is read by
    RAGVT5 (src/RAGVT5.py:294).  The three config keys are REQUIRED, as in the reference."""

    def __init__(self, config: dict):
        self.compute_stats = config["compute_stats"]
        self.compute_stats_examples = config["compute_stats_examples"] and self.compute_stats
        self.n_stats_examples = config["n_stats_examples"]
        self.stats: Dict[str, Any] = {}
        self.stats_examples: Dict[str, Dict[Any, list]] = {}

    def stat_sum(self, stat: str, key: Any, value: int = 1):
        if self.compute_stats:
            self.stats[stat][key] = self.stats[stat].get(key, 0) + value

    def stat_subtract(self, stat: str, key: Any, value: int = 1):
        return self.stat_sum(stat, key, -value)

    def stat_add_example(self, stat: str, key: Any, example: Any):
        if self.compute_stats_examples:
            bucket = self.stats_examples[stat].setdefault(key, [])
            if len(bucket) < self.n_stats_examples:
                bucket.append(example)

    def stat_remove_example(self, stat: str, key: Any, example: Any):
        if self.compute_stats_examples and key in self.stats_examples[stat]:
            try:
                self.stats_examples[stat][key].remove(example)
            except ValueError:
                pass


def _device_of(config: dict) -> torch.device:
    dev = torch.device(config.get("device", "cuda"))
    if dev.type != "cuda":
        raise RuntimeError("rag_docvqa_b200 runs on CUDA devices only (config['device']=%r)" % (config.get("device"),))
    if dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    return dev


def _to_device(t: torch.Tensor, dev: torch.device) -> torch.Tensor:
    return t if t.is_cuda else t.to(dev, non_blocking=True)


def compact_chunk(words: Sequence[str], boxes: Sequence[Sequence[float]]):
    """One chunk of Chunker.compact_chunks (src/_modules.py:1102-1132)."""
    if len(boxes):
        x0, y0, x1, y1 = zip(*boxes)
        return " ".join(words), [min(x0), min(y0), max(x1), max(y1)]
    return " ".join(words), [0, 0, 1, 1]


def crop_rectangle(bbox, width: int, height: int):
    """src/_modules.py:2108-2119: int() truncation, then order fix."""
    x0, y0, x1, y1 = int(bbox[0] * width), int(bbox[1] * height), int(bbox[2] * width), int(bbox[3] * height)
    return [min(x0, x1), min(y0, y1), max(x0, x1), max(y0, y1)]


def _lazy_crop_class():
    from PIL import Image

    class LazyCrop(Image.Image):
        """A PIL image whose pixels are cut out of `page` on first use -- the same deferred-load protocol
        PIL's own file images follow (every PIL operation calls load() before touching the pixels), so it
        behaves exactly like the eager `page.crop(rect)` of reference src/_modules.py:2119."""

        # Only PUBLIC Pillow protocol is used, so the class works with the reference's pinned Pillow 10.3 as well as with
        # Pillow >= 11: `mode` / `size` are plain attributes up to 10.0 and read-only properties over `_mode` / `_size`
        # from 10.1; `im` is a plain attribute up to 10.x and a property over `_im` from 11 -- assigning `self.im` is right
        # for both, and whether the pixels exist yet is this class's own state (`_page`), never Pillow's private field.
        _MODE_IS_PROPERTY = isinstance(getattr(Image.Image, "mode", None), property)
        _SIZE_IS_PROPERTY = isinstance(getattr(Image.Image, "size", None), property)

        def __init__(self, page, rect):
            super().__init__()                     # Pillow's own field set (it differs between Pillow versions)
            size = (rect[2] - rect[0], rect[3] - rect[1])
            if self._MODE_IS_PROPERTY:
                self._mode = page.mode
            else:
                self.mode = page.mode
            if self._SIZE_IS_PROPERTY:
                self._size = size
            else:
                self.size = size
            if page.info:
                self.info = dict(page.info)
            self._page, self._rect = page, rect

        def load(self):
            if self._page is not None:
                real = self._page.crop(self._rect)
                self.im, self.palette = real.im, real.palette
                self._page = None
            return super().load()
    return LazyCrop


_LazyCrop = None


def lazy_crop(page, rect):
    global _LazyCrop
    if _LazyCrop is None:
        _LazyCrop = _lazy_crop_class()
    return _LazyCrop(page, tuple(rect))


_CROP_POOL = None


def _crop_all(jobs):
    """page.crop(rect) for every (page, rect) job, in order.  PIL's crop is a row copy that releases the GIL, and
    it is the one part of the reference's output format that costs real time (src/_modules.py:2119: ~1.5 ms per
    patch, 63 % of the reference's retrieve), so the copies run on a small thread pool."""
    global _CROP_POOL
    if len(jobs) < 4:
        return [page.crop(rect) for page, rect in jobs]
    if _CROP_POOL is None:
        import os
        from concurrent.futures import ThreadPoolExecutor
        _CROP_POOL = ThreadPoolExecutor(max_workers=max(2, min(16, os.cpu_count() or 2)), thread_name_prefix="rdv-crop")
    return list(_CROP_POOL.map(lambda job: job[0].crop(job[1]), jobs))


def _defer_crop(page, rect):
    return (page, rect)


_SMALL_BATCH_BYTES = 1 << 20      # host batches up to this size take the one-upload / one-launch / one-read path


class Retriever(StatComponent):
    def __init__(self, config: dict):
        super().__init__(config)
        # optional keys (the defaults preserve the reference's behaviour).  retrieval_lazy_patches: cut the patch pixels
        # on first use.  retrieval_pause_gc: pause the cyclic garbage collector for the duration of a retrieve() call (a
        # process-wide switch: leave it off when other threads allocate).  retrieval_embedding_cache_mb: keep up to that
        # many MB of HOST document embeddings resident on the device, so that further questions about the same documents
        # do not cross PCIe again (0 = off).
        self.lazy_patches = bool(config.get("retrieval_lazy_patches", False))
        self.pause_gc = bool(config.get("retrieval_pause_gc", False))
        self._cache_budget = int(float(config.get("retrieval_embedding_cache_mb", 0)) * (1 << 20))
        self._cache = None          # functional.EmbeddingCache, built on first use (it needs the device)
        self.k = config.get("chunk_num", 10)
        self.include_surroundings = config.get("include_surroundings", 0)
        self.layout_map = get_layout_model_map(config)
        self.reorder_chunks = config.get("reorder_chunks", False)
        self.device = _device_of(config)
        self._small_bufs = {}
        if self.compute_stats:
            self.stats["layout_labels_topk_dist"] = {label: 0 for label in self.layout_map.values()}

    # -- a4 -----------------------------------------------------------------------------------------
    def _get_similarities(self, text_embeddings: List[torch.Tensor], question_embeddings: torch.Tensor):
        return self._score_topk(text_embeddings, question_embeddings).similarities

    def _score_topk(self, text_embeddings, question_embeddings) -> F.ScoreTopK:
        dev = question_embeddings.device if question_embeddings.is_cuda else self.device
        q = _to_device(question_embeddings, dev)
        if len(text_embeddings) and not any(e.is_cuda for e in text_embeddings):
            # host documents: one packed device buffer, the per-document copies issued from C
            if q.dim() != 2 or q.shape[0] != len(text_embeddings):
                raise ValueError("question_embeddings must be (B, d) with B == len(text_embeddings)")
            with torch.cuda.device(dev):
                table = F.upload_doc_table(text_embeddings, q.shape[1], dev)
                return F.score_topk_table(table, q, int(self.k))
        emb = [_to_device(e, dev) for e in text_embeddings]
        return F.score_topk(emb, q, int(self.k))

    # -- a7/a8/a9: host view of the hits ----------------------------------------------------------------
    def _hit_lists(self, hits: Sequence[Sequence[int]], words_text_chunks, words_box_chunks,
                   layout_labels_chunks, images, page_indices):
        s = self.include_surroundings
        bs = len(hits)
        if s == 0 and not self.reorder_chunks and _hostlists is not None:
            if self.lazy_patches:
                return _hostlists.gather_s0(hits, words_text_chunks, words_box_chunks, layout_labels_chunks, images,
                                            page_indices, lazy_crop)
            out = _hostlists.gather_s0(hits, words_text_chunks, words_box_chunks, layout_labels_chunks, images,
                                       page_indices, _defer_crop)
            crops = iter(_crop_all([job for doc in out[6] for job in doc]))
            for doc in out[6]:
                doc[:] = [next(crops) for _ in doc]
            return out
        out_text, out_bbox, out_labels, out_words, out_boxes, out_wlabels, out_patches, out_pages = (
            [], [], [], [], [], [], [], [])
        for b in range(bs):
            doc_hits = hits[b]
            labels = [layout_labels_chunks[b][i] for i in doc_hits]
            pages = [page_indices[b][i] for i in doc_hits]
            words_b, boxes_b = words_text_chunks[b], words_box_chunks[b]
            d_words, d_boxes = [], []
            if s == 0:
                # ranges of distinct chunks are disjoint in the page word list: the dedup is a no-op
                for i in doc_hits:
                    d_words.append(list(words_b[i]))
                    d_boxes.append(list(boxes_b[i]))
            else:
                page_arr = np.asarray(page_indices[b])
                page_view = {}     # page -> (chunk ids on the page, start offsets, page length)
                emitted = {}       # page -> list of raw [lo, hi) ranges of better hits
                for i, p in zip(doc_hits, pages):
                    if p not in page_view:
                        ids = np.nonzero(page_arr == p)[0]
                        lens = np.fromiter((len(words_b[c]) for c in ids), dtype=np.int64, count=len(ids))
                        starts = np.cumsum(lens) - lens
                        page_view[p] = (ids, starts, int(lens.sum()), lens)
                        emitted[p] = []
                    ids, starts, page_len, lens = page_view[p]
                    slot = int(np.searchsorted(ids, i))
                    start = int(starts[slot])
                    lo, hi = max(0, start - s), min(page_len, start + int(lens[slot]) + s)
                    fresh = [(lo, hi)]
                    for (cl, ch) in emitted[p]:
                        nxt = []
                        for (a, z) in fresh:
                            if ch <= a or cl >= z:
                                nxt.append((a, z))
                                continue
                            if a < cl:
                                nxt.append((a, cl))
                            if ch < z:
                                nxt.append((ch, z))
                        fresh = nxt
                    emitted[p].append((lo, hi))
                    w_out, b_out = [], []
                    for (a, z) in fresh:
                        first = int(np.searchsorted(starts + lens, a, side="right"))
                        for sl in range(first, len(ids)):
                            cs = int(starts[sl])
                            if cs >= z:
                                break
                            c = int(ids[sl])
                            x0, x1 = max(a, cs) - cs, min(z, cs + int(lens[sl])) - cs
                            if x0 < x1:
                                w_out.extend(words_b[c][x0:x1])
                                b_out.extend(boxes_b[c][x0:x1])
                    d_words.append(w_out)
                    d_boxes.append(b_out)
            texts, bboxes = [], []
            for w, bx in zip(d_words, d_boxes):
                t, bb = compact_chunk(w, bx)
                texts.append(t)
                bboxes.append(bb)
            wlabels = [[labels[j]] * len(d_words[j]) for j in range(len(d_words))]
            patches = []
            for j, p in enumerate(pages):
                page = images[b][p]
                rect = crop_rectangle(bboxes[j], page.width, page.height)
                patches.append(lazy_crop(page, rect) if self.lazy_patches else (page, tuple(rect)))
            if not self.lazy_patches:
                patches = _crop_all(patches)
            if self.reorder_chunks:
                order = sorted(range(len(pages)), key=lambda j: (pages[j], bboxes[j][1], bboxes[j][0]))
                texts = [texts[j] for j in order]
                bboxes = [bboxes[j] for j in order]
                labels = [labels[j] for j in order]
                d_words = [d_words[j] for j in order]
                d_boxes = [d_boxes[j] for j in order]
                wlabels = [wlabels[j] for j in order]
                patches = [patches[j] for j in order]
                pages = [pages[j] for j in order]
            out_text.append(texts); out_bbox.append(bboxes); out_labels.append(labels)
            out_words.append(d_words); out_boxes.append(d_boxes); out_wlabels.append(wlabels)
            out_patches.append(patches); out_pages.append(pages)
        return out_text, out_bbox, out_labels, out_words, out_boxes, out_wlabels, out_patches, out_pages

    def _get_top_k(self, similarities: List[torch.Tensor], words_text_chunks, words_box_chunks,
                   layout_labels_chunks, images, page_indices):
        """Same signature as the reference (src/_modules.py:1999): top-k of given similarity vectors."""
        dev = self.device
        sims = [_to_device(s_b, dev) for s_b in similarities]
        idx, _val, cnt = F.topk_segments(sims, int(self.k))
        hits = self._hits_to_host(idx, cnt)
        return self._hit_lists(hits, words_text_chunks, words_box_chunks, layout_labels_chunks, images, page_indices)

    @staticmethod
    def _hits_to_host(topk_idx: torch.Tensor, topk_cnt: torch.Tensor) -> List[List[int]]:
        packed = torch.cat([topk_idx, topk_cnt.unsqueeze(1)], dim=1).cpu().numpy()   # ONE D2H copy + sync
        k = topk_idx.shape[1]
        return [packed[b, :packed[b, k]].tolist() for b in range(packed.shape[0])]

    # -- a11 -----------------------------------------------------------------------------------------
    def retrieve(self, text_embeddings: List[torch.Tensor], question_embeddings: torch.Tensor,
                 words_text_chunks: list, words_box_chunks: list, layout_labels_chunks: list,
                 images: list, page_indices: list) -> tuple:
        """Retrieve the top-k chunks: 9-tuple, see reference src/_modules.py:2144-2153, 2180."""
        # The call builds ~1500 small acyclic containers next to the caller's millions of live word / box objects;
        # the cyclic collector's generation passes triggered by those allocations were up to 30 % of the call
        # (scripts/profile_e2e_phases.py).  Nothing created here can be part of a cycle, so with the opt-in key
        # `retrieval_pause_gc` collection is paused for the duration of the call and resumes (with its counters intact) on
        # return.  It is a process-wide switch, hence off by default.
        if not self.pause_gc:
            return self._retrieve(text_embeddings, question_embeddings, words_text_chunks, words_box_chunks,
                                  layout_labels_chunks, images, page_indices)
        gc_was_enabled = gc.isenabled()
        gc.disable()
        try:
            return self._retrieve(text_embeddings, question_embeddings, words_text_chunks, words_box_chunks,
                                  layout_labels_chunks, images, page_indices)
        finally:
            if gc_was_enabled:
                gc.enable()

    def _retrieve(self, text_embeddings, question_embeddings, words_text_chunks, words_box_chunks,
                  layout_labels_chunks, images, page_indices) -> tuple:
        inputs_on_host = not question_embeddings.is_cuda
        if self._cache_budget and inputs_on_host and len(text_embeddings) and not any(e.is_cuda for e in text_embeddings):
            # documents answered from the resident copies; the similarities still go back to the host, where the
            # caller's embeddings live (src/_modules.py:1990-1995)
            if self._cache is None:
                self._cache = F.EmbeddingCache(self._cache_budget, self.device)
            res = self._score_topk(self._cache.resident(text_embeddings), question_embeddings.to(self.device, non_blocking=True))
            hits = self._hits_to_host(res.topk_idx, res.topk_cnt)
            lists = self._hit_lists(hits, words_text_chunks, words_box_chunks, layout_labels_chunks, images, page_indices)
            return (*lists, list(torch.split(res.sims.cpu(), res.sizes)))
        if (inputs_on_host and len(text_embeddings) >= 8 and not any(e.is_cuda for e in text_embeddings)
                and question_embeddings.dim() == 2 and question_embeddings.shape[0] == len(text_embeddings)):
            return self._retrieve_host_pipelined(text_embeddings, question_embeddings, words_text_chunks, words_box_chunks,
                                                 layout_labels_chunks, images, page_indices)
        if (inputs_on_host and len(text_embeddings) and not any(e.is_cuda for e in text_embeddings)
                and question_embeddings.dim() == 2 and question_embeddings.shape[0] == len(text_embeddings)
                and question_embeddings.dtype == torch.float32
                and all(e.dtype == torch.float32 and e.dim() == 2 for e in text_embeddings)
                and (sum(e.shape[0] for e in text_embeddings) + len(text_embeddings)) * question_embeddings.shape[1] * 4
                <= _SMALL_BATCH_BYTES):
            return self._retrieve_host_small(text_embeddings, question_embeddings, words_text_chunks, words_box_chunks,
                                             layout_labels_chunks, images, page_indices)
        res = self._score_topk(text_embeddings, question_embeddings)
        hits = self._hits_to_host(res.topk_idx, res.topk_cnt)
        lists = self._hit_lists(hits, words_text_chunks, words_box_chunks, layout_labels_chunks, images, page_indices)
        sims = res.similarities
        if inputs_on_host:      # similarities live on the embeddings' device (src/_modules.py:1990-1995)
            flat = res.sims.cpu()
            sims = list(torch.split(flat, res.sizes))
        return (*lists, sims)

    def _small_buffers(self, dev, n_in: int, n_dev_out: int, n_host_out: int):
        """Pinned upload blob, its device twin, the device result buffer and its pinned read-back buffer, kept
        between calls (every call ends with a stream synchronise, so nothing is in flight when they are reused)."""
        c_in, c_dev, c_host = (max(2 * n, 1 << 16) for n in (n_in, n_dev_out, n_host_out))
        with torch.cuda.device(dev):
            bufs = (torch.empty(c_in, dtype=torch.uint8, pin_memory=True),
                    torch.empty(c_in, dtype=torch.uint8, device=dev),
                    torch.empty(c_dev, dtype=torch.uint8, device=dev),
                    torch.empty(c_host, dtype=torch.uint8, pin_memory=True))
            torch.cuda.current_stream(dev).synchronize()
        bufs = bufs + tuple(t.data_ptr() for t in bufs) + (bufs[3].numpy(),)
        self._small_bufs[dev.index] = bufs
        return bufs

    def _retrieve_host_small(self, text_embeddings, question_embeddings, words_text_chunks, words_box_chunks,
                             layout_labels_chunks, images, page_indices) -> tuple:
        """Host inputs of at most a megabyte (C1: one page of 30 chunks, the reference's own CPU-runnable case): a 46 KB
        problem is all fixed cost, so the device round trip is ONE C call (rdv_retrieve_small_f32): row offsets, tile
        descriptors, questions and embedding rows packed into one pinned blob, one upload, one launch (fused score +
        top-k: a cluster per document), one read-back of similarities, hits and counts, one synchronise.  Measured on B200 at C1: 0.164 ms per
        call through the general path (7 transfers, 2 launches) -> 0.124 ms with one blob each way driven from Python."""
        dev = self.device
        B, d, k = len(text_embeddings), int(question_embeddings.shape[1]), int(self.k)
        docs, sizes = [], []
        for b, e in enumerate(text_embeddings):
            n = e.shape[0]
            if n and e.shape[1] != d:
                raise ValueError("document %d: expected (n, %d) embeddings, got %s" % (b, d, tuple(e.shape)))
            docs.append(e.detach().contiguous())
            sizes.append(n)
        q = question_embeddings.detach().contiguous()
        h_docs = (ctypes.c_void_p * B)(*[t.data_ptr() if n else None for t, n in zip(docs, sizes)])
        rows = (ctypes.c_int64 * B)(*sizes)
        lay = F._lib.SmallLayoutStruct()
        bufs = self._small_bufs.get(dev.index) or self._small_buffers(dev, 0, 0, 0)
        stream = torch.cuda.current_stream(dev).cuda_stream
        call = F._lib_fn.rdv_retrieve_small_f32
        while True:
            host, blob, out, out_h, p_host, p_blob, p_out, p_out_h, res = bufs
            with torch.cuda.device(dev):
                rc = call(h_docs, rows, B, d, k, q.data_ptr(), p_host, p_blob, host.numel(), p_out, out.numel(), p_out_h,
                          out_h.numel(), ctypes.addressof(lay), stream)
            if rc != F._lib.SMALL_GROW:
                break
            bufs = self._small_buffers(dev, lay.in_bytes, lay.out_bytes, lay.read_bytes)
        if rc:
            F._lib.check(rc)
        o_idx, o_cnt = lay.o_idx, lay.o_cnt
        idx = res[o_idx:o_cnt].view(np.int32).reshape(B, k)
        cnt = res[o_cnt:lay.read_bytes].view(np.int32).tolist()
        hits = [idx[b, :cnt[b]].tolist() for b in range(B)]
        sims = torch.from_numpy(res[:o_idx].view(np.float32).copy())          # the caller owns its similarities
        lists = self._hit_lists(hits, words_text_chunks, words_box_chunks, layout_labels_chunks, images, page_indices)
        return (*lists, list(torch.split(sims, sizes)))

    def _retrieve_host_pipelined(self, text_embeddings, question_embeddings, words_text_chunks, words_box_chunks,
                                 layout_labels_chunks, images, page_indices) -> tuple:
        """Host inputs: the call is bounded by the H2D copy of the embeddings (~0.6 ms for a C2 batch) plus the list
        building (~1.3 ms).  The batch is cut into a few groups of documents, all enqueued at once; while the DMA
        engine still copies group g+1 the host already builds the lists of group g.  Same outputs, document order."""
        dev = self.device
        B, d, k = len(text_embeddings), question_embeddings.shape[1], int(self.k)
        lib, check = F._lib_fn, F._lib.check
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            q = F._f32_contig_aligned(question_embeddings.to(dev, non_blocking=True))
            table = F.upload_doc_table(text_embeddings, d, dev)          # ONE table for the batch
            sizes = np.asarray(table.sizes, dtype=np.int64)
            row_off = np.zeros(B + 1, dtype=np.int64)
            np.cumsum(sizes, out=row_off[1:])
            tile_off = np.zeros(B + 1, dtype=np.int64)
            np.cumsum((sizes + table.tile_rows - 1) // table.tile_rows, out=tile_off[1:])
            total = int(row_off[-1])
            n_groups = 4 if B >= 32 else 2
            # groups of ~equal row counts (the PCIe reads dominate)
            cuts = np.searchsorted(row_off[1:], total * np.arange(1, n_groups) / n_groups, side="left") + 1
            bounds = sorted(set([0, B] + [int(c) for c in cuts if 0 < c < B]))
            sims = torch.empty(max(total, 1), dtype=torch.float32, device=dev)
            idx = torch.empty((B, k), dtype=torch.int32, device=dev)
            val = torch.empty((B, k), dtype=torch.float32, device=dev)
            cnt = torch.empty((B,), dtype=torch.int32, device=dev)
            idx_h = torch.empty((B, k), dtype=torch.int32, pin_memory=True)
            cnt_h = torch.empty((B,), dtype=torch.int32, pin_memory=True)
            sims_h = torch.empty((total,), dtype=torch.float32, pin_memory=True)
            p_tiles, p_row = table.pointers()
            algo = table.algo
            events = []
            for lo, hi in zip(bounds[:-1], bounds[1:]):
                t_lo, t_hi, r_lo, r_hi = int(tile_off[lo]), int(tile_off[hi]), int(row_off[lo]), int(row_off[hi])
                if t_hi > t_lo:
                    check(lib.rdv_score_f32(p_tiles + 32 * t_lo, t_hi - t_lo, table.tile_rows, algo, q.data_ptr(), B, d,
                                            sims.data_ptr(), stream))
                check(lib.rdv_topk_segments_f32(sims.data_ptr(), p_row + 8 * lo, hi - lo, k, table.max_rows,
                                                idx.data_ptr() + 4 * k * lo, val.data_ptr() + 4 * k * lo,
                                                cnt.data_ptr() + 4 * lo, stream))
                idx_h[lo:hi].copy_(idx[lo:hi], non_blocking=True)
                cnt_h[lo:hi].copy_(cnt[lo:hi], non_blocking=True)
                if r_hi > r_lo:
                    sims_h[r_lo:r_hi].copy_(sims[r_lo:r_hi], non_blocking=True)
                done = torch.cuda.Event()
                done.record()
                events.append(done)
        outs = [[] for _ in range(8)]
        idx_np, cnt_np = idx_h.numpy(), cnt_h.numpy()
        for (lo, hi), done in zip(zip(bounds[:-1], bounds[1:]), events):
            done.synchronize()
            hits = [idx_np[b, :cnt_np[b]].tolist() for b in range(lo, hi)]
            lists = self._hit_lists(hits, words_text_chunks[lo:hi], words_box_chunks[lo:hi], layout_labels_chunks[lo:hi],
                                    images[lo:hi], page_indices[lo:hi])
            for o in range(8):
                outs[o].extend(lists[o])
        return (*outs, list(torch.split(sims_h, table.sizes)))

    def retrieve_packed(self, text_embeddings, question_embeddings, store, prompt_ids, sep_ids=(),
                        eos_id: int = 1, pad_id: int = 0, max_source_length: int = 512,
                        with_layout_labels: bool = False, pages=None, image_size: int = 224, resample: int = 3,
                        image_mean=(0.5, 0.5, 0.5), image_std=(0.5, 0.5, 0.5), return_plan: bool = False):
        """B200-native fast path: score -> top-k -> device gather straight into the generator's
        input_ids / boxes / attention_mask (what flatten + VT5.prepare_inputs_for_vqa build on the host,
        src/utils.py:233-253, src/VT5.py:141-185) for a pre-tokenised `DocStore`.  No Python lists.
        With `pages` (a PageStore) the retrieved patches are also cropped, grid-packed and resized on the device
        (page.crop + concatenate_patches(mode="grid") + the feature extractor's resize, src/_modules.py:2102-2121,
        src/utils.py:180-231, src/_modules.py:133); returns (packed, res, visual) in that case.  `return_plan` appends
        the GatherPlan, which postproc.Reranker.rerank_packed re-launches in the reranked order."""
        dev = question_embeddings.device if question_embeddings.is_cuda else self.device
        on_host = len(text_embeddings) and not any(e.is_cuda for e in text_embeddings)
        emb = [] if on_host else [_to_device(e, dev) for e in text_embeddings]
        q = _to_device(question_embeddings, dev)
        k = int(self.k)
        with torch.cuda.device(dev):
            if len(text_embeddings) and not any(e.is_cuda for e in text_embeddings):
                table = F.upload_doc_table(text_embeddings, q.shape[1], dev)
            else:
                table = F.build_doc_table(emb, q.shape[1], dev)
            B = table.B
            q = F._f32_contig_aligned(q)
            sims = torch.empty(table.total_rows, dtype=torch.float32, device=dev)
            topk_idx = torch.empty((B, k), dtype=torch.int32, device=dev)
            topk_val = torch.empty((B, k), dtype=torch.float32, device=dev)
            topk_cnt = torch.empty((B,), dtype=torch.int32, device=dev)
            plan = store.prepare_gather(topk_idx, topk_cnt, prompt_ids, include_surroundings=self.include_surroundings,
                                        reorder_chunks=self.reorder_chunks, sep_ids=sep_ids, eos_id=eos_id,
                                        pad_id=pad_id, max_len=max_source_length,
                                        with_layout_labels=with_layout_labels, sims=sims, topk_val=topk_val,
                                        max_rows=table.max_rows)
            if plan.prefers_one_launch(table):
                # ONE launch (thread-block clusters; csrc/retrieve_cluster.cu): only where rdv_retrieve_plan picks it
                plan.launch_retrieve(table, q, sims)
            else:
                # two launches: streaming score kernel, then one block per document selects its top-k and gathers
                F.score_table(table, q, out=sims)
                plan.launch()
            visual = None
            if pages is not None:
                visual = pages.pack(plan.t["hit_i"][1], plan.t["hit_rect"], topk_cnt, out_size=image_size, resample=resample,
                                    mean=image_mean, std=image_std)
            packed = plan.finish()
        res = F.ScoreTopK(list(torch.split(sims, table.sizes)) if B else [], sims, topk_idx, topk_val, topk_cnt,
                          table.sizes)
        out = (packed, res) if pages is None else (packed, res, visual)
        return out + (plan,) if return_plan else out


# ====================================================================================================
# visual path
# ====================================================================================================
def _surrounding_cells(row: int, col: int, n_rows: int, n_cols: int, include_surroundings):
    """Neighbourhood pattern of reference src/_modules.py:2207-2282."""
    cells = set()
    if isinstance(include_surroundings, (tuple, list)) and len(include_surroundings) == 2:
        rx, ry = include_surroundings
        for r in range(row - ry, row + ry + 1):
            for c in range(col - rx, col + rx + 1):
                cells.add((r, c))
    else:
        level, phase = divmod(int(include_surroundings), 3)
        for r in range(row - level, row + level + 1):
            for c in range(col - level, col + level + 1):
                cells.add((r, c))
            if phase > 0:
                cells.add((r, col - level - 1))
                cells.add((r, col + level + 1))
        if phase > 1:
            for c in range(col - level, col + level + 1):
                cells.add((row - level - 1, c))
                cells.add((row + level + 1, c))
    return [(r, c) for (r, c) in cells if 0 <= r < n_rows and 0 <= c < n_cols]


def _merge_rectangles(rects: List[Sequence[float]]) -> List[List[float]]:
    """Bounding box of every connected component of the strict-overlap graph
    (src/_modules.py:2331-2375, overlap test src/utils.py:460-463)."""
    n = len(rects)
    parent = list(range(n))

    def find(i):
        while parent[i] != i:
            parent[i] = parent[parent[i]]
            i = parent[i]
        return i
    for i in range(n):
        a = rects[i]
        for j in range(i + 1, n):
            b = rects[j]
            if a[0] < b[2] and a[2] > b[0] and a[1] < b[3] and a[3] > b[1]:
                parent[find(i)] = find(j)
    comps: Dict[int, list] = {}
    for i in range(n):
        comps.setdefault(find(i), []).append(rects[i])
    return sorted([min(r[0] for r in comp), min(r[1] for r in comp), max(r[2] for r in comp),
                   max(r[3] for r in comp)] for comp in comps.values())


class VisualRetriever:
    def __init__(self, config: dict, score: Optional[str] = None):
        self.k = config.get("chunk_num", 10)
        self.include_surroundings = config.get("include_surroundings", 0)
        self.mode = config.get("chunk_mode", "horizontal")
        self.layout_map = get_layout_model_map(config)
        self.device = _device_of(config)
        # optional key / argument (the default is the reference's MaxSim late interaction): "pooled" scores a strip by the
        # best cosine of its patch vectors against the mean-pooled question (functional.pooled_patch_topk)
        self.score = score or config.get("visual_score", "maxsim")
        if self.score not in ("maxsim", "pooled"):
            raise ValueError("VisualRetriever: visual_score must be 'maxsim' or 'pooled', got %r" % (self.score,))

    def _get_similarities(self, patch_embeddings: List[torch.Tensor], question_embeddings: torch.Tensor):
        """MaxSim of question i against the strips of document i (src/_modules.py:2191-2205).  Documents alternate
        between two side streams: the HBM-bound normalise + split of one document overlaps the tensor-bound
        contraction of the previous one."""
        dev = question_embeddings.device if question_embeddings.is_cuda else self.device
        q = _to_device(question_embeddings, dev)
        n_docs = len(patch_embeddings)
        if self.score == "pooled":
            return F.pooled_patch_topk([_to_device(p, dev) for p in patch_embeddings], q, int(self.k)).strip_scores
        if n_docs < 2:
            return [F.late_interaction(q[i].unsqueeze(0), _to_device(patch_embeddings[i], dev)) for i in range(n_docs)]
        if getattr(self, "_side_streams", None) is None or self._side_streams[0].device != dev:
            self._side_streams = [torch.cuda.Stream(dev) for _ in range(2)]
        cur = torch.cuda.current_stream(dev)
        ready = cur.record_event()
        sims = []
        for i in range(n_docs):
            st = self._side_streams[i & 1]
            st.wait_event(ready)
            with torch.cuda.stream(st):
                s_i = F.late_interaction(q[i].unsqueeze(0), _to_device(patch_embeddings[i], dev))
            s_i.record_stream(cur)
            sims.append(s_i)
        for st in self._side_streams:
            cur.wait_stream(st)
        return sims

    def _get_surrounding_patches(self, patch_coord, patches_matrix, include_surroundings=0):
        n_rows = len(patches_matrix)
        n_cols = len(patches_matrix[0]) if n_rows > 0 else 0
        return _surrounding_cells(patch_coord[0], patch_coord[1], n_rows, n_cols, include_surroundings)

    def _decode_rects(self, similarities, patches_flatten_indices, patches_matrix_list, patches_xyxy):
        """Top-k on the device, then the reference's integer decode on the host (src/_modules.py:2399-2448): per
        document the list of (group, merged rectangle) in sorted order, and the sorted group ids."""
        bs = len(similarities)
        # a document without strips still carries a (1,) dummy score (ImageEncoder returns zeros(1,2048,768),
        # src/_modules.py:1663-1664): it must not contribute hits  (src/_modules.py:2403-2406)
        idx, _val, cnt = F.topk_segments([_to_device(s_b, self.device) for s_b in similarities], int(self.k))
        hits = Retriever._hits_to_host(idx, cnt)
        rects_all, pages_all = [], []
        for b in range(bs):
            flat = np.asarray(patches_flatten_indices[b])
            if len(flat) == 0:
                rects_all.append([])
                pages_all.append([])
                continue
            cells = set()
            for i in hits[b]:
                group = int(flat[i])
                row = int(i - np.count_nonzero(flat < group))      # src/_modules.py:2411-2412
                if self.mode == "square":
                    raise NotImplementedError()                    # src/_modules.py:2413-2414
                matrix = patches_matrix_list[b][group]
                for (r, c) in self._get_surrounding_patches((row, 0), matrix, self.include_surroundings):
                    cells.add((group, r, c))
            by_group: Dict[int, list] = {}
            for (g, r, c) in cells:
                by_group.setdefault(g, []).append(list(patches_xyxy[b][g][r]))
            rects = []
            for g in sorted(by_group):
                for rect in _merge_rectangles(by_group[g]):
                    rects.append((g, tuple(rect)))
            rects_all.append(rects)
            pages_all.append(sorted(int(g) for g in by_group))
        return rects_all, pages_all

    def _get_top_k(self, similarities, patches_flatten_indices, patches_matrix_list, patches_xyxy, images):
        if len(similarities) == 0:
            return [], []
        rects_all, pages_all = self._decode_rects(similarities, patches_flatten_indices, patches_matrix_list, patches_xyxy)
        crops_all = [[images[b][g].crop(rect) for (g, rect) in rects] for b, rects in enumerate(rects_all)]   # src/_modules.py:2381
        return crops_all, pages_all

    def retrieve_packed(self, patch_embeddings, question_embeddings, patches_flatten_indices, patches_matrix_list,
                        patches_xyxy, pages, max_total_patches: int = 2048, patch: int = 16, normalize: bool = True):
        """B200-native fast path of the visual route: MaxSim + top-k on the device, the integer decode on the host, and
        the retrieved crops turned into the Pix2Struct generator's flattened patches on the device from a `PageStore`
        (what src/custom_pix2struct_processor.py:97-132, 175-196, 225 build from PIL crops; the header text of
        render_header is not drawn).  Returns (Pix2StructInputs, page ids); documents without a hit are not allowed,
        as in the reference (extract_multi_image_flattened_patches raises on an empty list)."""
        similarities = self._get_similarities(patch_embeddings, question_embeddings)
        rects_all, pages_all = self._decode_rects(similarities, patches_flatten_indices, patches_matrix_list, patches_xyxy)
        crops = [[(g,) + tuple(int(round(v)) for v in rect) for (g, rect) in rects] for rects in rects_all]   # PIL crop rounds its box
        return pages.pack_pix2struct(crops, max_total_patches=max_total_patches, patch=patch, normalize=normalize), pages_all

    def retrieve(self, patch_embeddings: List[torch.Tensor], question_embeddings: torch.Tensor,
                 patches_flatten_indices: list, patches_matrix_list: list, patches_xyxy: list,
                 images: List[list]) -> tuple:
        similarities = self._get_similarities(patch_embeddings, question_embeddings)
        return self._get_top_k(similarities, patches_flatten_indices, patches_matrix_list, patches_xyxy, images)
