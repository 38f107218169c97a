"""Seeded synthetic inputs shaped like the workloads BASELINE.json names.

Nothing here is reference code: the reference ships no data generator.  The shapes follow
the batch dict the reference's retrieval stage consumes (src/RAGVT5.py:208-252):

  text_embeddings        List[B] of (n_b, d) fp32      BiEncoder.batch_forward  (src/_modules.py:1415)
  question_embeddings    (B, d) fp32                   BiEncoder.forward        (src/_modules.py:1457)
  words_text_chunks      [B][n_b][w] str               Chunker.get_chunks       (src/_modules.py:1095)
  words_box_chunks       [B][n_b][w][4] float 0..1
  layout_labels_chunks   [B][n_b] int
  page_indices           [B][n_b] int
  images                 [B][pages] PIL.Image

Workload ids: C1..C5 are BASELINE.json `configs[0..4]`.
"""
from __future__ import annotations

import dataclasses
from typing import Dict, List, Optional

import numpy as np
import torch

SEED_BASE = 1234


@dataclasses.dataclass(frozen=True)
class Workload:
    name: str
    config_id: int
    docs: int            # B: one question per document
    min_pages: int
    max_pages: int
    chunks_per_page: int
    dim: int
    k: int


WORKLOADS: Dict[str, Workload] = {
    # configs[0]: SP-DocVQA-shaped, the reference's own CPU-runnable case
    "C1": Workload("C1", 1, 1, 1, 1, 30, 384, 5),
    # configs[1]: MP-DocVQA-shaped batch (the bench headline)
    "C2": Workload("C2", 2, 64, 1, 20, 30, 384, 5),
    # configs[2]: DUDE / MMLongBenchDoc-shaped long documents
    "C3": Workload("C3", 3, 256, 20, 200, 50, 768, 10),
}


def _gen(seed: int, device="cpu") -> torch.Generator:
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    return g


def doc_sizes(w: Workload, seed: Optional[int] = None, ragged_edge_cases: bool = True,
              full: bool = False) -> List[int]:
    """Number of chunks per document.  `full=True` gives every doc max_pages pages (the
    algorithmic-bytes figure in BASELINE.md assumes this)."""
    seed = SEED_BASE + w.config_id if seed is None else seed
    rng = np.random.RandomState(seed)
    if full:
        pages = np.full(w.docs, w.max_pages)
    else:
        pages = rng.randint(w.min_pages, w.max_pages + 1, size=w.docs)
    sizes = (pages * w.chunks_per_page).astype(np.int64)
    if ragged_edge_cases and w.docs >= 8 and not full:
        sizes[3] = 0            # an empty document  (src/_modules.py:1463-1464)
        sizes[5] = max(1, w.k - 2)  # fewer chunks than k  (src/_modules.py:2015)
    return [int(s) for s in sizes]


def make_embeddings(sizes: List[int], dim: int, seed: int, device="cpu", normalised: bool = True,
                    dup_frac: float = 0.01, dtype=torch.float32):
    """Chunk and question embeddings with a shared mean direction (cosines cluster ~0.2-0.6,
    like real sentence embedders) and `dup_frac` exact duplicate rows per document (repeated
    headers/footers), which exercise the lowest-index tie-break."""
    g = _gen(seed, device)
    B = len(sizes)
    u = torch.randn(dim, generator=g, device=device)
    u = u / u.norm()
    docs = []
    for n in sizes:
        e = torch.randn(n, dim, generator=g, device=device) / (dim ** 0.5) + 0.5 * u
        if normalised and n:
            e = e / e.norm(dim=-1, keepdim=True)
        n_dup = int(n * dup_frac)
        if n_dup:
            src = torch.randint(0, n, (n_dup,), generator=g, device=device)
            dst = torch.randint(0, n, (n_dup,), generator=g, device=device)
            e[dst] = e[src].clone()
        docs.append(e.to(dtype).contiguous())
    q = torch.randn(B, dim, generator=g, device=device) / (dim ** 0.5) + 0.5 * u
    if normalised:
        q = q / q.norm(dim=-1, keepdim=True)
    return docs, q.to(dtype).contiguous()


def make_page_indices(sizes: List[int], chunks_per_page: int) -> List[List[int]]:
    return [[i // chunks_per_page for i in range(n)] for n in sizes]


def make_words(sizes: List[int], seed: int, min_words: int = 40, max_words: int = 72,
               vocab: int = 5000, empty_chunk_every: int = 0):
    """Per-chunk word strings and boxes (plain Python lists, the reference's format).
    Returns (words_text_chunks, words_box_chunks, layout_labels_chunks)."""
    rng = np.random.RandomState(seed)
    words_all, boxes_all, labels_all = [], [], []
    for n in sizes:
        words_doc, boxes_doc = [], []
        for c in range(n):
            w = int(rng.randint(min_words, max_words + 1))
            if empty_chunk_every and c % empty_chunk_every == empty_chunk_every - 1:
                w = 0
            ids = rng.randint(0, vocab, size=w)
            x0 = rng.uniform(0.0, 0.9, size=w)
            y0 = rng.uniform(0.0, 0.95, size=w)
            x1 = x0 + rng.uniform(0.005, 0.1, size=w)
            y1 = y0 + rng.uniform(0.005, 0.05, size=w)
            words_doc.append(["w%d" % i for i in ids])
            boxes_doc.append(np.stack([x0, y0, x1, y1], axis=1).tolist() if w else [])
        words_all.append(words_doc)
        boxes_all.append(boxes_doc)
        labels_all.append([1 + (c % 3 == 0) * 2 for c in range(n)])  # labels in {1, 3}
    return words_all, boxes_all, labels_all


def make_images(sizes: List[int], chunks_per_page: int, width: int = 850, height: int = 1100,
                ragged_sizes: bool = False, share_pool: int = 0):
    """One blank-ish RGB page per page index (PIL).  Content is a cheap gradient so crops differ.
    share_pool > 0: documents share a pool of that many distinct page images (bench memory saver;
    every document still indexes its own page list)."""
    from PIL import Image
    out = []
    pool = {}
    for b, n in enumerate(sizes):
        n_pages = max(1, -(-n // chunks_per_page))
        pages = []
        for p in range(n_pages):
            if share_pool:
                key = (b * 7 + p) % share_pool
                if key in pool:
                    pages.append(pool[key])
                    continue
            w = width + (17 * p if ragged_sizes else 0)
            h = height - (13 * p if ragged_sizes else 0)
            arr = np.empty((h, w, 3), dtype=np.uint8)
            arr[..., 0] = (np.arange(w, dtype=np.uint32) * 255 // max(1, w - 1)).astype(np.uint8)[None, :]
            arr[..., 1] = (np.arange(h, dtype=np.uint32) * 255 // max(1, h - 1)).astype(np.uint8)[:, None]
            arr[..., 2] = (b * 37 + p * 11) % 256
            pages.append(Image.fromarray(arr, "RGB"))
            if share_pool:
                pool[(b * 7 + p) % share_pool] = pages[-1]
        out.append(pages)
    return out


def make_text_batch(name: str, with_lists: bool = False, device="cpu", normalised: bool = True,
                    seed: Optional[int] = None, docs: Optional[int] = None, full: bool = False,
                    dup_frac: float = 0.01, ragged_edge_cases: bool = True, with_images: bool = True,
                    share_image_pool: int = 0, emb_seed: Optional[int] = None):
    """One batch of workload `name` (C1/C2/C3).  `docs` overrides B (for slices); `emb_seed` reseeds the
    embedding VALUES only (same document sizes / lists) -- used to rotate distinct resident batches."""
    w = WORKLOADS[name]
    if docs is not None:
        w = dataclasses.replace(w, docs=docs)
    seed = SEED_BASE + w.config_id if seed is None else seed
    sizes = doc_sizes(w, seed, ragged_edge_cases=ragged_edge_cases, full=full)
    emb, q = make_embeddings(sizes, w.dim, seed if emb_seed is None else emb_seed, device=device,
                             normalised=normalised, dup_frac=dup_frac)
    batch = dict(workload=w, sizes=sizes, text_embeddings=emb, question_embeddings=q,
                 page_indices=make_page_indices(sizes, w.chunks_per_page))
    if with_lists:
        words, boxes, labels = make_words(sizes, seed + 7)
        batch.update(words_text_chunks=words, words_box_chunks=boxes, layout_labels_chunks=labels)
        if with_images:
            batch["images"] = make_images(sizes, w.chunks_per_page, share_pool=share_image_pool)
    return batch


def make_token_batch(n: int, dim: int, seed: int, device="cpu", mean_len: float = 96.0,
                     std_len: float = 24.0, min_len: int = 8, max_len: int = 512,
                     all_pad_rows: int = 0):
    """Token embeddings + int64 right-padded attention mask, the inputs of mean_pooling
    (src/_model_utils.py:49; produced at src/_modules.py:1466-1473)."""
    g = _gen(seed, device)
    lens = torch.clamp(torch.round(torch.randn(n, generator=g, device=device) * std_len + mean_len),
                       min_len, max_len).to(torch.int64)
    if all_pad_rows:
        lens[:all_pad_rows] = 0
    L = int(lens.max().item()) if n else 0
    L = max(L, 1)
    embs = torch.randn(n, L, dim, generator=g, device=device)
    mask = (torch.arange(L, device=device)[None, :] < lens[:, None]).to(torch.int64)
    return embs, mask


def make_strip_batch(docs: int, strips: List[int], tokens: int, dim: int, seed: int, device="cpu"):
    """Un-pooled encoder token matrices for the visual path: List[B] of (n_b, tokens, dim) and
    questions (B, tokens, dim)  (src/_modules.py:1627-1666 produces (n, 2048, 768))."""
    g = _gen(seed, device)
    patches = [torch.randn(n, tokens, dim, generator=g, device=device) for n in strips]
    q = torch.randn(docs, tokens, dim, generator=g, device=device)
    return patches, q


def make_tokens_for_words(words_text_chunks, seed: int = 0, vocab: int = 32000):
    """Deterministic stand-in for the T5 tokenizer (no tokenizer files exist offline): each
    distinct word maps to 1-3 ids (p = .6/.3/.1), ids in [3, vocab).  Returns a dict word->list."""
    table = {}
    rng = np.random.RandomState(seed)
    for doc in words_text_chunks:
        for chunk in doc:
            for w in chunk:
                if w not in table:
                    n_tok = int(rng.choice([1, 2, 3], p=[0.6, 0.3, 0.1]))
                    table[w] = [int(t) for t in rng.randint(3, vocab, size=n_tok)]
    return table


def make_chunker_batch(seed: int, docs: int = 4, max_pages: int = 6, max_words: int = 400, max_layouts: int = 14,
                       clusters: bool = False, degenerate: bool = True, numpy_pages: bool = False):
    """Chunker.get_chunks inputs in the reference's format (src/_modules.py:872-898): words [B][pages][n] str, boxes
    [B][pages][n][4] float 0..1, layout_info [B][pages] {"boxes": [n_l][4], "labels": [n_l] (, "clusters": ndarray)}.
    Words are laid out on text lines; layout boxes are rectangles over the page, some overlapping, some cutting words
    exactly in half (ratio == 0.5 is NOT inside), plus pages without layout boxes / without words, zero-area words and
    duplicate (xmin, ymin) keys for the stable sort."""
    rng = np.random.RandomState(seed)
    words, boxes, info = [], [], []
    for b in range(docs):
        n_pages = int(rng.randint(1, max_pages + 1))
        d_words, d_boxes, d_info = [], [], []
        for p in range(n_pages):
            n = int(rng.randint(0, max_words + 1)) if (degenerate and rng.rand() < 0.15) else int(rng.randint(max_words // 4, max_words + 1))
            if degenerate and rng.rand() < 0.05:
                n = 0
            per_line = int(rng.randint(6, 16))
            line = np.arange(n) // per_line
            col = np.arange(n) % per_line
            # a 1/1024 grid keeps coordinates exactly representable, so exact halves (ratio == 0.5) really occur
            x0 = (32 + col * 60 + rng.randint(0, 8, size=n)) / 1024.0
            y0 = (40 + line * 24) / 1024.0 % 0.95
            x1 = x0 + rng.choice([16, 32, 48, 56], size=n) / 1024.0
            y1 = y0 + 16 / 1024.0
            if degenerate and n:
                z = rng.rand(n) < 0.02
                x1 = np.where(z, x0, x1)                      # zero-area word: ratio 0 by definition
            pb = np.stack([x0, y0, x1, y1], axis=1)
            d_words.append(["w%d" % i for i in rng.randint(0, 5000, size=n)])
            d_boxes.append(pb if numpy_pages else pb.tolist())
            n_l = 0 if (degenerate and rng.rand() < 0.2) else int(rng.randint(1, max_layouts + 1))
            lx0 = rng.randint(0, 700, size=n_l) // 8 * 8
            ly0 = rng.randint(0, 800, size=n_l) // 8 * 8
            lx1 = lx0 + rng.randint(64, 600, size=n_l) // 8 * 8
            ly1 = ly0 + rng.randint(24, 400, size=n_l) // 8 * 8
            if degenerate and n_l > 2:
                lx0[1], ly0[1] = lx0[0], ly0[0]               # equal sort keys: order must stay as given
            lb = (np.stack([lx0, ly0, lx1, ly1], axis=1) / 1024.0).tolist()
            page = {"boxes": lb, "labels": [int(x) for x in rng.randint(0, 11, size=n_l)]}
            if clusters:
                page["clusters"] = rng.choice([-1, -1, 0, 1, 2], size=n_l).astype(np.int64)
            d_info.append(page)
        words.append(d_words); boxes.append(d_boxes); info.append(d_info)
    return words, boxes, info


def make_s2_pages(seed: int, pages: int = 6, max_layouts: int = 30, max_words: int = 300, degenerate: bool = True):
    """S2Chunker.forward inputs (reference src/_modules.py:1929-1962): layout_info [pages] {"boxes": [n_l][4] float 0..1,
    "labels": [n_l]} and pages_info [pages] {"ocr_tokens": [n] str, "ocr_normalized_boxes": [n][4]}.  Layout boxes have
    arbitrary float64 coordinates (pixel / page size, as LayoutModel produces them, :520-524), some pages have no or one
    layout box, some boxes coincide (distance 0) and some hold no word."""
    rng = np.random.RandomState(seed)
    layout_info, pages_info = [], []
    for p in range(pages):
        n_l = int(rng.randint(2, max_layouts + 1))
        if degenerate and p % 5 == 3:
            n_l = int(rng.randint(0, 2))
        w, h = int(rng.randint(600, 2600)), int(rng.randint(800, 3400))
        x0 = rng.randint(0, w - 50, size=n_l); y0 = rng.randint(0, h - 50, size=n_l)
        x1 = np.minimum(w, x0 + rng.randint(20, w // 2, size=n_l)); y1 = np.minimum(h, y0 + rng.randint(10, h // 3, size=n_l))
        boxes = [[int(a) / w, int(b) / h, int(c) / w, int(d) / h] for a, b, c, d in zip(x0, y0, x1, y1)]
        if degenerate and n_l > 3:
            boxes[2] = list(boxes[0])                                    # coincident regions: distance 0, weight 1
        layout_info.append({"boxes": boxes, "labels": [int(x) for x in rng.randint(0, 11, size=n_l)]})
        n = int(rng.randint(max_words // 4, max_words + 1))
        wx0 = rng.uniform(0.0, 0.92, size=n); wy0 = rng.uniform(0.0, 0.96, size=n)
        wb = np.stack([wx0, wy0, wx0 + rng.uniform(0.01, 0.07, size=n), wy0 + rng.uniform(0.008, 0.02, size=n)], axis=1)
        pages_info.append({"ocr_tokens": ["w%d" % i for i in rng.randint(0, 5000, size=n)],
                           "ocr_normalized_boxes": wb.tolist()})
    return layout_info, pages_info


class HashEmbedder:
    """Stand-in for BiEncoder (a producer outside the path): a deterministic (n, dim) fp32 embedding per text, with a
    shared direction so cosines are mostly positive; an empty list gives (0, dim) as the reference does (:1463-1464)."""

    def __init__(self, dim: int = 384, device="cpu"):
        self.dim, self.device = dim, device
        self.common = torch.randn(dim, generator=_gen(99))

    def forward(self, texts: List[str]) -> torch.Tensor:
        import zlib
        rows = [torch.randn(self.dim, generator=_gen(zlib.crc32(t.encode()) & 0x7FFFFFFF)) + 0.7 * self.common for t in texts]
        out = torch.stack(rows) if rows else torch.empty(0, self.dim)
        return out.to(self.device)
