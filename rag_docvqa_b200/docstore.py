"""Pre-tokenised CSR document store + the device gather into the generator's input tensors.

Host side: turn the reference's nested lists (words_text_chunks / words_box_chunks /
layout_labels_chunks / page_indices, src/RAGVT5.py:208-224) into flat CSR arrays ONCE per batch of
documents, upload them in one pinned copy.  Device side: rdv_gather_vt5_inputs (csrc/gather.cu) builds
input_ids / boxes / attention_mask / layout labels from the top-k kernel's output without a host round
trip (reference: src/_modules.py:2014-2142, src/utils.py:233-253, src/VT5.py:141-185).
"""
from __future__ import annotations

import ctypes
from typing import Callable, List, NamedTuple, Optional, Sequence

import numpy as np
import torch

from . import _lib
from .functional import _stream_ptr

try:                                   # C helper for the per-word walk (csrc/hostlists.c); the Python walk below is the
    from . import _hostlists           # same algorithm and stays as the specification
except ImportError:                    # pragma: no cover - built by rag_docvqa_b200.build
    _hostlists = None

_I32 = np.int32


class PackedInputs(NamedTuple):
    input_ids: torch.Tensor        # (B, longest) int64
    boxes: torch.Tensor            # (B, longest, 4) int64
    attention_mask: torch.Tensor   # (B, longest) int64
    layout_labels: Optional[torch.Tensor]
    longest: int
    hit_chunk: torch.Tensor        # (B, k) int32, OUTPUT order (after reorder_chunks), -1 padded
    hit_page: torch.Tensor
    hit_label: torch.Tensor
    hit_nwords: torch.Tensor
    hit_bbox: torch.Tensor         # (B, k, 4) float64
    hit_rect: torch.Tensor         # (B, k, 4) int32 crop rectangle (x0, y0, x1, y1), -1 without page sizes


class DocStore:
    """CSR view of a batch of documents on the device (see `rdv_docstore` in include/rdv.h)."""

    FIELDS = ("chunk_rec", "chunk_off", "chunk_word_off", "word_tok_off", "tok_ids", "tok_word", "word_box", "chunk_label", "chunk_page",
              "chunk_page_start", "page_chunks", "run_begin", "run_end", "doc_page_off", "page_wh", "tok_rec", "chunk_bbox")

    def __init__(self, arrays: dict, B: int, device):
        self.B = B
        self.device = device
        self.host = arrays
        # one pinned blob, one H2D copy
        offsets, total = {}, 0
        for name in self.FIELDS:
            arr = arrays.get(name)
            if arr is None:
                continue
            total = (total + 15) // 16 * 16
            offsets[name] = total
            total += arr.nbytes
        blob = torch.empty(max(total, 16), dtype=torch.uint8, pin_memory=True)
        raw = blob.numpy()
        for name, off in offsets.items():
            arr = arrays[name]
            raw[off:off + arr.nbytes] = np.frombuffer(arr.tobytes(), dtype=np.uint8)
        self.blob = blob.to(device, non_blocking=True)
        base = self.blob.data_ptr()
        self.struct = _lib.DocStoreStruct()
        self.struct.B = B
        for name in self.FIELDS:
            setattr(self.struct, name, base + offsets[name] if name in offsets else None)
        self.n_chunks = int(arrays["chunk_off"][-1])
        self.n_words = int(arrays["chunk_word_off"][-1])
        self.n_tokens = int(arrays["word_tok_off"][-1])

    def same_chunks_as(self, sizes) -> bool:
        """True when a batch of embeddings with `sizes` rows per document describes exactly this store's chunks."""
        sizes = np.asarray(sizes, dtype=np.int64)
        return len(sizes) == self.B and bool(np.array_equal(np.diff(self.host["chunk_off"]), sizes))

    @classmethod
    def from_lists(cls, words_text_chunks, words_box_chunks, layout_labels_chunks, page_indices,
                   tokenize: Callable[[str], Sequence[int]], device, images=None, derived: bool = True) -> "DocStore":
        """tokenize(word) -> the word's token ids WITHOUT the trailing EOS (src/VT5.py:160).
        derived=False leaves out the optional gather-friendly views (tok_rec, chunk_bbox); the kernel then
        walks tok_ids -> tok_word -> word_box (tests cover both)."""
        arrays, B = cls.arrays_from_lists(words_text_chunks, words_box_chunks, layout_labels_chunks, page_indices, tokenize,
                                          images=images, derived=derived)
        return cls(arrays, B, device)

    @staticmethod
    def arrays_from_lists(words_text_chunks, words_box_chunks, layout_labels_chunks, page_indices, tokenize,
                          images=None, derived: bool = True):
        """The host half of from_lists: nested lists -> CSR arrays (no device needed)."""
        B = len(words_text_chunks)
        sizes = np.array([len(doc) for doc in words_text_chunks], dtype=np.int64)
        chunk_off = np.zeros(B + 1, dtype=np.int64)
        np.cumsum(sizes, out=chunk_off[1:])
        N = int(chunk_off[-1])
        chunk_nwords = np.zeros(N, dtype=np.int64)
        chunk_label = np.zeros(N, dtype=_I32)
        chunk_page = np.zeros(N, dtype=_I32)
        g = 0
        for b in range(B):
            labels_b, pages_b = layout_labels_chunks[b], page_indices[b]
            for c in range(int(sizes[b])):
                chunk_label[g] = labels_b[c]
                chunk_page[g] = pages_b[c]
                g += 1
        tok_cache = {}
        if _hostlists is not None:
            # the walk over every word (2 M for a C2 batch) in C: csrc/hostlists.c flatten_docs
            raw_nw, raw_nt, raw_ids, raw_boxes = _hostlists.flatten_docs(words_text_chunks, words_box_chunks, tokenize, tok_cache)
            chunk_nwords = np.frombuffer(raw_nw, dtype=np.int64)
            word_ntok = np.frombuffer(raw_nt, dtype=_I32)
            tok_ids = np.frombuffer(raw_ids, dtype=_I32)
            boxes = np.frombuffer(raw_boxes, dtype=np.float64)
            if len(chunk_nwords) != N:
                raise ValueError("words_text_chunks: %d chunks, expected %d" % (len(chunk_nwords), N))
        else:
            word_ntok, tok_ids, boxes = [], [], []
            g = 0
            for b in range(B):
                for c in range(int(sizes[b])):
                    words = words_text_chunks[b][c]
                    chunk_nwords[g] = len(words)
                    for w in words:
                        toks = tok_cache.get(w)
                        if toks is None:
                            toks = tok_cache[w] = tuple(int(t) for t in tokenize(w))
                        word_ntok.append(len(toks))
                        tok_ids.extend(toks)
                    if len(words):
                        if len(words_box_chunks[b][c]) != len(words):
                            raise ValueError("document %d chunk %d: %d words but %d boxes" % (b, c, len(words), len(words_box_chunks[b][c])))
                        boxes.extend(words_box_chunks[b][c])
                    g += 1
        chunk_word_off = np.zeros(N + 1, dtype=np.int64)
        np.cumsum(chunk_nwords, out=chunk_word_off[1:])
        W = int(chunk_word_off[-1])
        word_tok_off = np.zeros(W + 1, dtype=np.int64)
        np.cumsum(np.asarray(word_ntok, dtype=np.int64), out=word_tok_off[1:])
        word_box = np.asarray(boxes, dtype=np.float64).reshape(W, 4) if W else np.zeros((0, 4), dtype=np.float64)
        # page runs: chunks grouped by (document, page), chunk order inside a page (src/_modules.py:2032-2050)
        doc_of = np.repeat(np.arange(B, dtype=np.int64), sizes)
        order = np.lexsort((np.arange(N), chunk_page.astype(np.int64), doc_of)) if N else np.zeros(0, dtype=np.int64)
        key = doc_of[order] * (1 << 32) + chunk_page[order].astype(np.int64) if N else np.zeros(0, dtype=np.int64)
        new_run = np.ones(N, dtype=bool)
        if N > 1:
            new_run[1:] = key[1:] != key[:-1]
        run_id = np.cumsum(new_run) - 1 if N else np.zeros(0, dtype=np.int64)
        run_starts = np.nonzero(new_run)[0] if N else np.zeros(0, dtype=np.int64)
        run_ends = np.append(run_starts[1:], N) if N else np.zeros(0, dtype=np.int64)
        nw_sorted = chunk_nwords[order] if N else np.zeros(0, dtype=np.int64)
        csum = np.cumsum(nw_sorted) - nw_sorted if N else np.zeros(0, dtype=np.int64)
        start_sorted = csum - csum[run_starts][run_id] if N else np.zeros(0, dtype=np.int64)
        chunk_page_start = np.zeros(N, dtype=_I32)
        run_begin = np.zeros(N, dtype=_I32)
        run_end = np.zeros(N, dtype=_I32)
        if N:
            chunk_page_start[order] = start_sorted
            run_begin[order] = run_starts[run_id]
            run_end[order] = run_ends[run_id]
        rec = np.zeros((N, 8), dtype=_I32)                       # rdv_chunk_rec
        if N:
            rec[:, 0] = chunk_word_off[:-1]; rec[:, 1] = chunk_word_off[1:]
            rec[:, 2] = word_tok_off[chunk_word_off[:-1]]; rec[:, 3] = word_tok_off[chunk_word_off[1:]]
            rec[:, 4] = chunk_page; rec[:, 5] = chunk_label; rec[:, 6] = chunk_page_start
        tok_ids_a = np.asarray(tok_ids, dtype=_I32)
        tok_word = np.repeat(np.arange(W, dtype=_I32), np.asarray(word_ntok, dtype=np.int64)) if W else np.zeros(0, dtype=_I32)
        # derived, gather-friendly views (include/rdv.h: rdv_tok_rec, chunk_bbox)
        views = {}
        box1000 = word_box * 1000.0                               # the float64 product of src/VT5.py:162
        if derived and (W == 0 or bool(np.all(np.abs(box1000) < 2.0 ** 31))):   # NaN/inf/huge boxes: keep the f64 path
            tok_rec = np.zeros((len(tok_ids_a), 8), dtype=_I32)
            tok_rec[:, 0] = tok_ids_a
            tok_rec[:, 1] = tok_word
            if W:
                tok_rec[:, 2:6] = box1000.astype(np.int64)[tok_word]      # truncation toward zero, as int()
            views["tok_rec"] = tok_rec
        chunk_bbox = np.tile(np.array([0.0, 0.0, 1.0, 1.0]), (N, 1))       # empty chunk: src/_modules.py:1126-1127
        nonempty = chunk_nwords > 0
        if W:
            starts = chunk_word_off[:-1][nonempty]
            chunk_bbox[nonempty, :2] = np.minimum.reduceat(word_box[:, :2], starts, axis=0)
            chunk_bbox[nonempty, 2:] = np.maximum.reduceat(word_box[:, 2:], starts, axis=0)
        if derived:
            views["chunk_bbox"] = np.ascontiguousarray(chunk_bbox)
        arrays = dict(
            **views,
            chunk_rec=np.ascontiguousarray(rec), chunk_off=chunk_off, chunk_word_off=chunk_word_off.astype(_I32), word_tok_off=word_tok_off.astype(_I32),
            tok_ids=tok_ids_a, tok_word=tok_word,
            word_box=np.ascontiguousarray(word_box),
            chunk_label=chunk_label, chunk_page=chunk_page, chunk_page_start=chunk_page_start,
            page_chunks=order.astype(_I32), run_begin=run_begin, run_end=run_end)
        if images is not None:
            n_pages = np.array([len(p) for p in images], dtype=np.int64)
            doc_page_off = np.zeros(B + 1, dtype=np.int64)
            np.cumsum(n_pages, out=doc_page_off[1:])
            wh = np.array([[im.width, im.height] for pages in images for im in pages], dtype=_I32).reshape(-1, 2)
            arrays["doc_page_off"] = doc_page_off.astype(_I32)
            arrays["page_wh"] = np.ascontiguousarray(wh)
        return arrays, B

    # ---- on-disk form: the CSR arrays of a pre-tokenised batch of documents --------
    def save(self, path: str) -> None:
        """One .npz with every host array; DocStore.load(path, device) restores an identical store."""
        np.savez(path, __B=np.int64(self.B), **{k: v for k, v in self.host.items() if v is not None})

    @classmethod
    def load(cls, path: str, device) -> "DocStore":
        with np.load(path if path.endswith(".npz") else path + ".npz") as z:
            arrays = {k: z[k] for k in z.files if k != "__B"}
            B = int(z["__B"])
        return cls(arrays, B, device)

    # ------------------------------------------------------------------------------------------
    def prepare_gather(self, topk_idx: torch.Tensor, topk_cnt: torch.Tensor, prompt_ids: Sequence[Sequence[int]],
                       include_surroundings: int = 0, reorder_chunks: bool = False, sep_ids: Sequence[int] = (),
                       eos_id: int = 1, pad_id: int = 0, max_len: int = 512, with_layout_labels: bool = False,
                       max_seg: int = 32, sims: Optional[torch.Tensor] = None, topk_val: Optional[torch.Tensor] = None,
                       max_rows: int = 0, emit_order: Optional[torch.Tensor] = None,
                       emit_cnt: Optional[torch.Tensor] = None) -> "GatherPlan":
        """Allocates the outputs and fills the argument block of rdv_gather_vt5_inputs (no launch).
        With `sims` (all similarities of the batch, chunk order) the kernel selects the top-k itself and
        writes topk_idx / topk_val / topk_cnt: a step is then score kernel + this kernel."""
        dev = self.device
        B, k = topk_idx.shape
        if B != self.B:
            raise ValueError("gather: topk_idx has %d documents, store has %d" % (B, self.B))
        if isinstance(include_surroundings, (tuple, list)):
            raise ValueError("tuple include_surroundings is a VisualRetriever option")
        p_off = np.zeros(B + 1, dtype=_I32)
        np.cumsum([len(p) for p in prompt_ids], out=p_off[1:])
        flat = np.fromiter((t for p in prompt_ids for t in p), dtype=_I32, count=int(p_off[-1]))
        sep = np.asarray(list(sep_ids), dtype=_I32)
        host = torch.empty(p_off.nbytes + flat.nbytes + sep.nbytes + 16, dtype=torch.uint8, pin_memory=True)
        raw = host.numpy()
        raw[:p_off.nbytes] = np.frombuffer(p_off.tobytes(), dtype=np.uint8)
        raw[p_off.nbytes:p_off.nbytes + flat.nbytes] = np.frombuffer(flat.tobytes(), dtype=np.uint8)
        o_sep = p_off.nbytes + flat.nbytes
        raw[o_sep:o_sep + sep.nbytes] = np.frombuffer(sep.tobytes(), dtype=np.uint8)
        small = host.to(dev, non_blocking=True)

        i32 = dict(dtype=torch.int32, device=dev)
        t = dict(
            small=small, topk_idx=topk_idx, topk_cnt=topk_cnt,
            out_ids=torch.empty((B, max_len), dtype=torch.int64, device=dev),
            out_boxes=torch.empty((B, max_len, 4), dtype=torch.int64, device=dev),
            out_mask=torch.empty((B, max_len), dtype=torch.int64, device=dev),
            out_labels=torch.empty((B, max_len), dtype=torch.int64, device=dev) if with_layout_labels else None,
            meta=torch.empty((2, B), **i32),                  # full_len, status
            hit_i=torch.empty((4, B, k), **i32),              # chunk, page, label, nwords
            hit_bbox=torch.empty((B, k, 4), dtype=torch.float64, device=dev),
            hit_rect=torch.empty((B, k, 4), **i32),
            seg_ws=torch.empty((B * k * max_seg * 2,), **i32))
        a = _lib.GatherArgsStruct()
        a.topk_idx = topk_idx.data_ptr(); a.topk_cnt = topk_cnt.data_ptr()
        a.k = k; a.include_surroundings = int(include_surroundings); a.reorder_chunks = 1 if reorder_chunks else 0
        a.n_sep = len(sep)
        a.prompt_off = small.data_ptr(); a.prompt_ids = small.data_ptr() + p_off.nbytes
        a.sep_ids = small.data_ptr() + o_sep
        a.eos_id = eos_id; a.pad_id = pad_id; a.max_len = max_len; a.max_seg = max_seg
        a.seg_ws = t["seg_ws"].data_ptr()
        a.out_ids = t["out_ids"].data_ptr(); a.out_boxes = t["out_boxes"].data_ptr()
        a.out_mask = t["out_mask"].data_ptr()
        a.out_labels = t["out_labels"].data_ptr() if t["out_labels"] is not None else None
        a.full_len = t["meta"][0].data_ptr(); a.status = t["meta"][1].data_ptr()
        a.hit_chunk = t["hit_i"][0].data_ptr(); a.hit_page = t["hit_i"][1].data_ptr()
        a.hit_label = t["hit_i"][2].data_ptr(); a.hit_nwords = t["hit_i"][3].data_ptr()
        a.hit_bbox = t["hit_bbox"].data_ptr(); a.hit_rect = t["hit_rect"].data_ptr()
        if sims is not None:
            if topk_val is None:
                raise ValueError("fused selection needs topk_val")
            if sims.numel() != self.n_chunks:
                raise ValueError("gather: %d similarities for a store of %d chunks (the embeddings and the store must "
                                 "describe the same chunks)" % (sims.numel(), self.n_chunks))
            t["sims"], t["topk_val"] = sims, topk_val
            a.sims = sims.data_ptr(); a.topk_val = topk_val.data_ptr(); a.max_rows = int(max_rows)
        plan = GatherPlan(self, a, t, max_len, max_seg)
        if emit_order is not None:
            plan.set_emit_order(emit_order, emit_cnt)
        return plan

    def gather(self, topk_idx: torch.Tensor, topk_cnt: torch.Tensor, prompt_ids: Sequence[Sequence[int]],
               **options) -> PackedInputs:
        """prepare + launch rdv_gather_vt5_inputs on the current stream + ONE small D2H read."""
        plan = self.prepare_gather(topk_idx, topk_cnt, prompt_ids, **options)
        plan.launch()
        return plan.finish()


class GatherPlan:
    """A filled argument block of rdv_gather_vt5_inputs: launch() enqueues the kernel (no sync),
    finish() reads full_len/status back (the step's one device->host read) and slices the outputs."""

    def __init__(self, store: DocStore, args, tensors: dict, max_len: int, max_seg: int):
        self.store, self.args, self.t, self.max_len, self.max_seg = store, args, tensors, max_len, max_seg
        self._ds_ref = ctypes.byref(store.struct)
        self._args_ref = ctypes.byref(args)

    def set_emit_order(self, order: Optional[torch.Tensor], cnt: Optional[torch.Tensor]) -> None:
        """The reranker's index list (postproc.rerank_order): (B, k) int32 positions of retrieve()'s output order and
        the number kept per document; None = everything in retrieval order.  The next launch() rebuilds the packed
        tensors and the hit_* arrays IN PLACE in that order."""
        if order is None:
            self.t.pop("emit_order", None); self.t.pop("emit_cnt", None)
            self.args.emit_order = None; self.args.emit_cnt = None
            return
        B, k = self.t["topk_idx"].shape
        if cnt is None or tuple(order.shape) != (B, k) or tuple(cnt.shape) != (B,):
            raise ValueError("emit_order must be (B, k) with emit_cnt (B,)")
        if order.dtype != torch.int32 or cnt.dtype != torch.int32 or not order.is_cuda or not cnt.is_cuda:
            raise ValueError("emit_order / emit_cnt must be int32 CUDA tensors")
        self.t["emit_order"], self.t["emit_cnt"] = order.contiguous(), cnt.contiguous()
        self.args.emit_order = self.t["emit_order"].data_ptr(); self.args.emit_cnt = self.t["emit_cnt"].data_ptr()

    def launch(self, stream: Optional[int] = None) -> None:
        dev = self.store.device
        rc = _lib.lib.rdv_gather_vt5_inputs(self._ds_ref, self._args_ref,
                                            _stream_ptr(dev) if stream is None else stream)
        if rc:
            _lib.check(rc)

    def can_retrieve_in_one_launch(self, table) -> bool:
        """rdv_retrieve_vt5_f32's requirements (include/rdv.h): the cluster kernel's limits, no neighbour windows, no
        reranked re-emission, and a store that describes the same chunks as the embeddings."""
        a = self.args
        return (a.include_surroundings == 0 and not a.emit_order and table.B == self.store.B and "topk_val" in self.t
                and table.cluster_fits(a.k) and self.store.same_chunks_as(table.sizes))

    def prefers_one_launch(self, table) -> bool:
        """... and rdv_retrieve_plan picks it over the streaming kernel + the select / gather kernel."""
        return self.can_retrieve_in_one_launch(table) and table.use_cluster(self.args.k)

    def launch_retrieve(self, table, questions: torch.Tensor, sims: torch.Tensor, stream: Optional[int] = None) -> None:
        """The whole step in ONE launch (rdv_retrieve_vt5_f32): cosine scores of `table` against `questions` into `sims`,
        per-document top-k into this plan's topk_idx / topk_val / topk_cnt, and the gather of the hits."""
        dev = self.store.device
        d_ctas, n_ctas, cluster = table.cluster_pointers()
        rc = _lib.lib.rdv_retrieve_vt5_f32(d_ctas, n_ctas, cluster, questions.data_ptr(), table.d, table.max_rows, sims.data_ptr(),
                                           self._ds_ref, self._args_ref, _stream_ptr(dev) if stream is None else stream)
        if rc:
            _lib.check(rc)

    def finish(self) -> PackedInputs:
        t = self.t
        meta_h = t["meta"].cpu()
        B = meta_h.shape[1]
        if B and int(meta_h[1].max()):
            raise _lib.RdvError(_lib.E_LIMIT, "gather: segment workspace overflow, raise max_seg (%d)" % self.max_seg)
        longest = min(int(meta_h[0].max()), self.max_len) if B else 0
        labels = t["out_labels"]
        return PackedInputs(t["out_ids"][:, :longest], t["out_boxes"][:, :longest], t["out_mask"][:, :longest],
                            labels[:, :longest] if labels is not None else None, longest,
                            t["hit_i"][0], t["hit_i"][1], t["hit_i"][2], t["hit_i"][3], t["hit_bbox"], t["hit_rect"])
