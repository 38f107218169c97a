"""Chunker with the word -> layout-box assignment on the device.

Drop-in for src._modules.Chunker (:843-1100): same constructor keys (`chunk_size`, `chunk_size_tol`, `overlap`,
`page_retrieval`, `cluster_layouts`, the three stat keys), same `get_chunks(words, boxes, layout_info, **kwargs)`
5-tuple and the same counters.  What the reference spends its time on with a layout model --
`containment_ratio(word, layout box) > 0.5` for every word x layout box of every page in Python (:1023-1033) -- is ONE
kernel launch for the whole batch (`rdv_layout_assign`: float64, bit-exact decisions); the host keeps what is list
surgery: the stable (xmin, ymin) order of a page's few layout boxes, cluster grouping and the chunk windows, which are
computed as index ranges instead of repeated list extension.
"""
from __future__ import annotations

from collections import Counter
from typing import List, Optional

import numpy as np
import torch

from . import _lib
from .functional import _stream_ptr
from .retriever import StatComponent, _device_of, get_layout_model_map


def chunk_ranges(n: int, chunk_size: int, overlap: int, tol: float):
    """Chunks of a list of n words as (begin, end, size_as_counted) ranges -- what the window loop of
    src/_modules.py:906-938 builds: windows of `chunk_size` every `chunk_size - overlap` words; a window joins the
    previous chunk while the reference's running size (prev + len(window) - overlap, which undercounts once a window is
    shorter than the overlap) stays within chunk_size * (1 + tol).  Also yields the (size, delta) counter events."""
    ranges, events = [], []
    prev = 0
    limit = chunk_size * (1 + tol)
    for i in range(0, n, chunk_size - overlap):
        end = min(n, i + chunk_size)
        size = end - i
        if i > 0 and prev + (size - overlap) <= limit:
            size = prev + size - overlap
            begin = ranges[-1][0]
            ranges[-1] = (begin, max(ranges[-1][1], end))
            events.append((prev, -1)); events.append((size, 1))
        else:
            ranges.append((i, end))
            events.append((size, 1))
        prev = size
    return ranges, events


class Chunker(StatComponent):
    def __init__(self, config: dict):
        super().__init__(config)
        self.chunk_size = config.get("chunk_size", 60)
        self.chunk_size_tol = config.get("chunk_size_tol", 0.2)
        self.overlap = config.get("overlap", 10)
        self.page_retrieval = config.get("page_retrieval", "concat")
        self.default_layout_label = {v: k for k, v in get_layout_model_map(config).items()}["text"]
        self.cluster_layouts = config.get("cluster_layouts", False)
        self.device = _device_of(config)
        if self.compute_stats:
            self.stats = {"chunk_size_dist": Counter(), "n_chunks_per_page_dist": Counter(),
                          "n_chunks_per_doc_dist": Counter()}
            if not config["layout_model_weights"] or self.page_retrieval == "oracle":
                self.stats["n_chunks_per_layout_dist"] = Counter()
        if self.compute_stats_examples:
            self.stats_examples = {key: {} for key in self.stats}
        assert self.chunk_size > 1, "chunk_size should be a non-negative non-zero integer."
        assert 0 <= self.chunk_size_tol <= 1, "chunk_size_tol should be a float between 0 and 1."
        assert self.overlap >= 0, "overlap should be a non-negative integer."
        assert self.overlap < self.chunk_size, "overlap should be less than chunk_size."

    # ------------------------------------------------------------------------------------------
    @staticmethod
    def compact_chunks(words_text_chunks: list, words_boxes_chunks: list) -> tuple:
        """Static on the class, as in the reference (src/_modules.py:1102-1132): src/RAGVT5.py:224 calls
        `Chunker.compact_chunks(...)` on the NAME, so a drop-in class must carry it.  [B][n][w] words / boxes ->
        [B][n] joined text and [B][n][4] bounding box ([0, 0, 1, 1] for a chunk without words)."""
        from .retriever import compact_chunk
        text_chunks, boxes_chunks = [], []
        for doc_words, doc_boxes in zip(words_text_chunks, words_boxes_chunks):
            pairs = [compact_chunk(w, bx) for w, bx in zip(doc_words, doc_boxes)]
            text_chunks.append([t for t, _ in pairs])
            boxes_chunks.append([bb for _, bb in pairs])
        return text_chunks, boxes_chunks

    def assign_words_to_layouts(self, page_boxes: List[np.ndarray], layout_boxes: List[np.ndarray],
                                layout_labels: List[np.ndarray], default_label: int = -1):
        """One launch for a list of pages.  page_boxes[i] (n_i,4) float64 word boxes, layout_boxes[i] (l_i,4) float64 in
        visiting order, layout_labels[i] (l_i,) int32 (get_chunks passes positions, so labels may be any object).
        Returns (inside, word_labels): inside[i] is a (l_i, n_i) bool array, word_labels[i] an (n_i,) int32 array
        holding the label of the last containing box, else `default_label`."""
        P = len(page_boxes)
        n_w = np.asarray([len(x) for x in page_boxes], dtype=np.int64)
        n_l = np.asarray([len(x) for x in layout_boxes], dtype=np.int64)
        groups = (n_w + 31) // 32
        page_word_off = np.zeros(P + 1, dtype=np.int32); np.cumsum(n_w, out=page_word_off[1:])
        page_lay_off = np.zeros(P + 1, dtype=np.int32); np.cumsum(n_l, out=page_lay_off[1:])
        page_group_off = np.zeros(P + 1, dtype=np.int32); np.cumsum(groups, out=page_group_off[1:])
        n_groups, G, W = int(page_group_off[-1]), int(page_lay_off[-1]), int(page_word_off[-1])
        if n_groups == 0 or G == 0:
            return ([np.zeros((int(n_l[i]), int(n_w[i])), dtype=bool) for i in range(P)],
                    [np.full(int(n_w[i]), default_label, dtype=np.int32) for i in range(P)])
        group_page = np.repeat(np.arange(P, dtype=np.int32), groups)
        bits_off = np.zeros(G + 1, dtype=np.int64)
        np.cumsum(np.repeat(groups, n_l), out=bits_off[1:])
        word_box = np.concatenate([np.asarray(x, dtype=np.float64).reshape(-1, 4) for x in page_boxes])
        lay_box = np.concatenate([np.asarray(x, dtype=np.float64).reshape(-1, 4) for x in layout_boxes])
        lay_label = np.concatenate([np.asarray(x, dtype=np.int32).reshape(-1) for x in layout_labels])
        # one pinned blob, one H2D copy (every array padded to 16 bytes: the boxes are read as double2)
        parts = [word_box, lay_box, bits_off, page_word_off, page_lay_off, page_group_off, group_page, lay_label]
        offs, total = [], 0
        for a in parts:
            offs.append(total)
            total += (a.nbytes + 15) // 16 * 16
        host = torch.empty(max(total, 16), dtype=torch.uint8, pin_memory=True)
        raw = host.numpy()
        for a, o in zip(parts, offs):
            raw[o:o + a.nbytes] = np.frombuffer(a.tobytes(), dtype=np.uint8)
        dev = self.device
        with torch.cuda.device(dev):
            blob = host.to(dev, non_blocking=True)
            base = blob.data_ptr()
            bits = torch.empty(int(bits_off[-1]), dtype=torch.int32, device=dev)
            word_label = torch.empty(W, dtype=torch.int32, device=dev)
            ptr = [base + o for o in offs]
            _lib.check(_lib.lib.rdv_layout_assign(ptr[0], ptr[3], ptr[1], ptr[7], ptr[4], ptr[6], ptr[5], n_groups,
                                                  int(default_label), ptr[2], bits.data_ptr(),
                                                  word_label.data_ptr(), _stream_ptr(dev)))
            bits_h = bits.cpu().numpy().view(np.uint32)
            label_h = word_label.cpu().numpy()
        inside, labels = [], []
        for i in range(P):
            nl, nw, ng = int(n_l[i]), int(n_w[i]), int(groups[i])
            rows = bits_h[bits_off[page_lay_off[i]]:bits_off[page_lay_off[i]] + nl * ng].reshape(nl, ng)
            unpacked = np.unpackbits(rows.view(np.uint8), axis=1, bitorder="little")[:, :nw].astype(bool) if nl and ng else \
                np.zeros((nl, nw), dtype=bool)
            inside.append(unpacked)
            labels.append(label_h[page_word_off[i]:page_word_off[i + 1]])
        return inside, labels

    # ------------------------------------------------------------------------------------------
    def _emit(self, words, boxes, words_out, boxes_out, example_id) -> int:
        ranges, events = chunk_ranges(len(words), self.chunk_size, self.overlap, self.chunk_size_tol)
        for size, delta in events:
            self.stat_sum("chunk_size_dist", size, delta)
            if delta > 0:
                self.stat_add_example("chunk_size_dist", size, example_id)
            else:
                self.stat_remove_example("chunk_size_dist", size, example_id)
        for lo, hi in ranges:
            words_out.append(words[lo:hi]); boxes_out.append(boxes[lo:hi])
        return len(ranges)

    def get_chunks(self, words: list, boxes: list, layout_info: Optional[list] = None, **kwargs) -> tuple:
        bs = len(words)
        question_id = kwargs.get("question_id", None)
        use_layout = layout_info is not None and layout_info != [[]]      # [[]] = no layout model (src/_modules.py:892)
        use_clusters = use_layout and "clusters" in layout_info[0][0].keys() and self.cluster_layouts
        oracle = self.page_retrieval == "oracle"

        # pass 1: normalise the page boxes as the reference does (:957-960) and collect every page with layout boxes
        norm_boxes, jobs = [], {}
        pb_list, lb_list, ll_list = [], [], []
        for b in range(bs):
            doc = []
            for p, (page_words, page_boxes) in enumerate(zip(words[b], boxes[b])):
                if not isinstance(page_words, list):
                    page_boxes = page_boxes.tolist()
                if len(page_boxes) > 0 and not isinstance(page_boxes[0], list):
                    page_boxes = [pbox.tolist() for pbox in page_boxes]
                doc.append(page_boxes)
                if oracle or not use_layout or len(layout_info[b][p]["boxes"]) == 0:
                    continue
                lb = layout_info[b][p]["boxes"]
                order = sorted(range(len(lb)), key=lambda j: (lb[j][0], lb[j][1]))      # stable, as sorted(zip(...)) (:1006-1018)
                jobs[(b, p)] = (len(pb_list), order)
                pb_list.append(np.asarray(page_boxes, dtype=np.float64).reshape(-1, 4))
                lb_list.append(np.asarray([lb[j] for j in order], dtype=np.float64).reshape(-1, 4))
                ll_list.append(np.arange(len(order), dtype=np.int32))      # the kernel carries positions; labels stay objects
            norm_boxes.append(doc)
        inside_all, labels_all = self.assign_words_to_layouts(pb_list, lb_list, ll_list) if pb_list else ([], [])

        # pass 2: chunks
        layout_labels_chunks, page_indices, words_text_chunks, words_boxes_chunks, words_layout_labels_pages = [], [], [], [], []
        for b in range(bs):
            d_labels, d_pages, d_words, d_boxes, d_word_labels, d_n = [], [], [], [], [], 0
            for p, page_words in enumerate(words[b]):
                page_boxes = norm_boxes[b][p]
                ex = f"{question_id[b]}_p{p}" if question_id is not None else None
                if oracle:
                    d_pages.append(p); d_words.append(page_words); d_boxes.append(page_boxes)
                    d_labels.append(self.default_layout_label)
                    d_word_labels.append([self.default_layout_label] * len(page_words))
                    d_n += 1
                    self.stat_sum("chunk_size_dist", len(page_words)); self.stat_sum("n_chunks_per_page_dist", 1)
                    self.stat_add_example("chunk_size_dist", len(page_words), ex)
                    self.stat_add_example("n_chunks_per_page_dist", 1, ex)
                    continue
                if (b, p) not in jobs:
                    n = self._emit(page_words, page_boxes, d_words, d_boxes, ex)
                    d_pages.extend([p] * n)
                    d_labels.extend([self.default_layout_label] * n)
                    d_word_labels.append([self.default_layout_label] * len(page_words))
                    d_n += n
                    self.stat_sum("n_chunks_per_page_dist", n); self.stat_add_example("n_chunks_per_page_dist", n, ex)
                    continue
                slot, order = jobs[(b, p)]
                inside, word_labels = inside_all[slot], labels_all[slot]
                labels = box_labels = [layout_info[b][p]["labels"][j] for j in order]
                members = [np.flatnonzero(row).tolist() for row in inside]          # word indices per layout box
                if use_clusters:
                    clusters = layout_info[b][p]["clusters"].tolist()
                    clusters = [clusters[j] for j in order]
                    if clusters:                                                    # :1035 (`if page_layout_clusters:`)
                        grouped, votes, seen = [], [], {}
                        for idx, lab, c in zip(members, labels, clusters):
                            if c != -1 and c in seen:
                                grouped[seen[c]].extend(idx); votes[seen[c]][lab] += 1
                            else:
                                if c != -1:
                                    seen[c] = len(grouped)
                                grouped.append(list(idx)); votes.append(Counter([lab]))
                        members = grouped
                        labels = [v.most_common(1)[0][0] for v in votes]
                page_n, n_before = 0, len(d_words)
                for idx, lab in zip(members, labels):
                    n = self._emit([page_words[i] for i in idx], [page_boxes[i] for i in idx], d_words, d_boxes, ex)
                    page_n += n
                    d_labels.extend([lab] * n)
                    self.stat_sum("n_chunks_per_layout_dist", n); self.stat_add_example("n_chunks_per_layout_dist", n, ex)
                d_pages.extend([p] * (len(d_words) - n_before))
                d_word_labels.append([box_labels[j] if j >= 0 else self.default_layout_label for j in word_labels.tolist()])
                d_n += page_n
                self.stat_sum("n_chunks_per_page_dist", page_n); self.stat_add_example("n_chunks_per_page_dist", page_n, ex)
            layout_labels_chunks.append(d_labels); page_indices.append(d_pages); words_text_chunks.append(d_words)
            words_boxes_chunks.append(d_boxes); words_layout_labels_pages.append(d_word_labels)
            self.stat_sum("n_chunks_per_doc_dist", d_n)
            self.stat_add_example("n_chunks_per_doc_dist", d_n, f"{question_id[b]}" if question_id is not None else None)
        return words_text_chunks, words_boxes_chunks, layout_labels_chunks, page_indices, words_layout_labels_pages
