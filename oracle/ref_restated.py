"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's retrieval hot path.

This file is the *checker*.  Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline /
`--impl reference` legs may import it; the product (rag_docvqa_b200/) never does.

Parity status: PINNED.  The reference (Pikurrot/RAG-DocVQA) has no tests or golden vectors of its
own, so the pins were manufactured: oracle/make_golden.py imports the unmodified
reference in the build container (oracle/ref_import.py), runs it on seeded inputs and freezes its
outputs under tests/golden/; tests/test_oracle_golden.py checks every function below against those
files (bit-exact: same torch CPU ops in the same order), and -- when /root/reference is present --
against the live reference as well.  The components either side of the path are pinned the same way:
oracle/make_golden_postproc.py (the reference's Reranker.rerank and the page vote of RAGVT5.forward, run as
written on a stand-in `self`) -> postproc.json, tests/test_postproc_oracle.py; oracle/make_golden_chunker.py (the
reference's Chunker.get_chunks) -> chunker.json, tests/test_chunker_oracle.py; oracle/make_golden_s2chunker.py (the reference's
S2Chunker node building, weight matrices and forward) -> s2chunker.json, tests/test_s2chunker_oracle.py.

Each function cites the reference lines it restates.  Floating-point work uses the same torch CPU
operators the reference calls (torch.norm / matmul / topk / F.normalize / bmm), so results are
bit-identical to the reference on the same torch build; integer / list work is plain Python.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple, Union

import numpy as np
import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------------------
# a4: cosine score          reference: src/_modules.py:1978-1997 (Retriever._get_similarities)
# --------------------------------------------------------------------------------------
def score(text_embeddings: Sequence[torch.Tensor], question_embeddings: torch.Tensor) -> List[torch.Tensor]:
    out = []
    for b, emb in enumerate(text_embeddings):
        q = question_embeddings[b]
        chunk_norms = torch.norm(emb, dim=-1)            # :1990
        q_norm = torch.norm(q)                            # :1991
        dots = torch.matmul(emb, q)                       # :1992
        out.append(dots / (chunk_norms * q_norm + 1e-8))  # :1993  eps on the *product* of norms
    return out


# --------------------------------------------------------------------------------------
# a6: per-document top-k    reference: src/_modules.py:2015-2016, :2408 (torch.topk)
# --------------------------------------------------------------------------------------
def topk_reference(sim: torch.Tensor, k: int) -> torch.Tensor:
    """Exactly what the reference calls; tie order is whatever torch.topk gives."""
    k_min = min(k, len(sim))                              # :2015
    return torch.topk(sim, k=k_min, dim=-1).indices       # :2016


def order_key(values: np.ndarray) -> np.ndarray:
    """fp32 -> uint32 key whose unsigned order is the float order with torch.topk's conventions:
    NaN (either sign) greatest, -0.0 == +0.0."""
    v = np.ascontiguousarray(values, dtype=np.float32)
    u = v.view(np.uint32).copy()
    u[u == np.uint32(0x80000000)] = 0                     # -0.0 -> +0.0
    neg = (u >> 31).astype(bool)
    key = np.where(neg, ~u, u | np.uint32(0x80000000)).astype(np.uint32)
    key[np.isnan(v)] = np.uint32(0xFFFFFFFF)
    return key


def topk_lowest_index(sim: Union[torch.Tensor, np.ndarray], k: int) -> np.ndarray:
    """The contract north_star mandates: descending score, ties broken by LOWEST index.
    Equal to torch.topk wherever scores are distinct."""
    v = sim.detach().cpu().numpy() if isinstance(sim, torch.Tensor) else np.asarray(sim)
    n = v.shape[0]
    k_min = min(k, n)
    if k_min == 0:
        return np.zeros(0, dtype=np.int64)
    packed = (order_key(v).astype(np.uint64) << np.uint64(32)) | \
        (np.uint64(0xFFFFFFFF) - np.arange(n, dtype=np.uint64))
    order = np.argsort(packed, kind="stable")[::-1]
    return order[:k_min].astype(np.int64)


# --------------------------------------------------------------------------------------
# a8: compact chunks        reference: src/_modules.py:1102-1132 (Chunker.compact_chunks)
# --------------------------------------------------------------------------------------
def compact_chunks(words_text_chunks, words_boxes_chunks):
    texts, bboxes = [], []
    for doc_words, doc_boxes in zip(words_text_chunks, words_boxes_chunks):
        doc_texts, doc_bboxes = [], []
        for chunk_words, chunk_boxes in zip(doc_words, doc_boxes):
            doc_texts.append(" ".join(chunk_words))
            if len(chunk_boxes):
                bbox = [min(bx[0] for bx in chunk_boxes), min(bx[1] for bx in chunk_boxes),
                        max(bx[2] for bx in chunk_boxes), max(bx[3] for bx in chunk_boxes)]
            else:
                bbox = [0, 0, 1, 1]                       # :1126-1127
            doc_bboxes.append(bbox)
        texts.append(doc_texts)
        bboxes.append(doc_bboxes)
    return texts, bboxes


# --------------------------------------------------------------------------------------
# a9: crop rectangle        reference: src/_modules.py:2102-2121
# --------------------------------------------------------------------------------------
def crop_rectangle(bbox, page_width: int, page_height: int):
    x0 = int(bbox[0] * page_width)                        # int() truncates toward zero  :2110-2113
    y0 = int(bbox[1] * page_height)
    x1 = int(bbox[2] * page_width)
    y1 = int(bbox[3] * page_height)
    return [min(x0, x1), min(y0, y1), max(x0, x1), max(y0, y1)]  # :2115-2118


# --------------------------------------------------------------------------------------
# a7: gather of the hits    reference: src/_modules.py:1999-2153 (Retriever._get_top_k)
# --------------------------------------------------------------------------------------
def gather_hits(topk_indices: Sequence[Sequence[int]], words_text_chunks, words_box_chunks,
                layout_labels_chunks, images, page_indices, include_surroundings: int = 0,
                reorder_chunks: bool = False, crop: bool = True):
    """Given per-document hit indices in rank order, rebuild the eight list outputs.
    With crop=False the 7th output holds crop rectangles instead of PIL images."""
    B = len(topk_indices)
    s = include_surroundings
    out_words, out_boxes, out_labels, out_pages = [], [], [], []
    for b in range(B):
        hits = [int(i) for i in topk_indices[b]]
        out_labels.append([layout_labels_chunks[b][i] for i in hits])   # :2019
        out_pages.append([page_indices[b][i] for i in hits])            # :2020
        # per-page concatenation of ALL chunks in chunk order  (:2032-2050)
        page_words, page_boxes, span, seen = {}, {}, {}, {}
        for c in range(len(words_text_chunks[b])):
            p = page_indices[b][c]
            if p not in page_words:
                page_words[p], page_boxes[p], seen[p] = [], [], set()
            start = len(page_words[p])
            page_words[p].extend(words_text_chunks[b][c])
            page_boxes[p].extend(words_box_chunks[b][c])
            span[c] = (start, start + len(words_text_chunks[b][c]))
        doc_words, doc_boxes = [], []
        for i in hits:                                                  # :2056-2087
            p = page_indices[b][i]
            start, end = span[i]
            lo = max(0, start - s)
            hi = min(len(page_words[p]), end + s)
            fresh = [j for j in range(lo, hi) if j not in seen[p]]     # dedup vs higher-ranked hits
            seen[p].update(fresh)
            doc_words.append([page_words[p][j] for j in fresh])
            doc_boxes.append([page_boxes[p][j] for j in fresh])
        out_words.append(doc_words)
        out_boxes.append(doc_boxes)
    out_text, out_bbox = compact_chunks(out_words, out_boxes)           # :2093
    out_word_labels = [[[out_labels[b][i]] * len(out_words[b][i]) for i in range(len(out_words[b]))]
                       for b in range(B)]                               # :2094-2100
    out_patches = []
    for b in range(B):                                                  # :2102-2121
        doc_patches = []
        for i, p in enumerate(out_pages[b]):
            page = images[b][p]
            rect = crop_rectangle(out_bbox[b][i], page.width, page.height)
            doc_patches.append(page.crop(rect) if crop else rect)
        out_patches.append(doc_patches)
    if reorder_chunks:                                                  # :2129-2142
        for b in range(B):
            order = sorted(range(len(out_pages[b])),
                           key=lambda i: (out_pages[b][i], out_bbox[b][i][1], out_bbox[b][i][0]))
            for lst in (out_text, out_bbox, out_labels, out_words, out_boxes, out_word_labels,
                        out_patches, out_pages):
                lst[b] = [lst[b][i] for i in order]
    return (out_text, out_bbox, out_labels, out_words, out_boxes, out_word_labels, out_patches, out_pages)


# --------------------------------------------------------------------------------------
# a11: Retriever.retrieve   reference: src/_modules.py:2155-2180
# --------------------------------------------------------------------------------------
def retrieve(text_embeddings, question_embeddings, words_text_chunks, words_box_chunks,
             layout_labels_chunks, images, page_indices, k: int = 10, include_surroundings: int = 0,
             reorder_chunks: bool = False, deterministic_ties: bool = False, crop: bool = True):
    sims = score(text_embeddings, question_embeddings)
    if deterministic_ties:
        hits = [topk_lowest_index(s_b, k) for s_b in sims]
    else:
        hits = [topk_reference(s_b, k).tolist() for s_b in sims]
    lists = gather_hits(hits, words_text_chunks, words_box_chunks, layout_labels_chunks, images,
                        page_indices, include_surroundings, reorder_chunks, crop=crop)
    return (*lists, sims)


# --------------------------------------------------------------------------------------
# a1: masked mean pooling   reference: src/_model_utils.py:49-61
# --------------------------------------------------------------------------------------
def mean_pooling(embs: torch.Tensor, attention_mask: torch.Tensor) -> torch.Tensor:
    expanded = attention_mask.unsqueeze(-1).expand(embs.size())         # :56
    summed = (embs * expanded).sum(dim=1)                               # :57-58
    counts = attention_mask.sum(dim=1).unsqueeze(-1).clamp(min=1e-9)    # :59
    return summed / counts                                              # :60


# --------------------------------------------------------------------------------------
# a5: MaxSim late interaction   reference: src/utils.py:442-458, src/_modules.py:2191-2205
# --------------------------------------------------------------------------------------
def pooled_patch_scores(patch_embeddings: Sequence[torch.Tensor], question_embeddings: torch.Tensor,
                        question_mask: torch.Tensor = None):
    """Pooled-patch visual retrieval, composed from the reference's own functions: the question tokens through
    mean_pooling (src/_model_utils.py:49-61), every patch vector of every strip through Retriever._get_similarities
    (src/_modules.py:1978-1997), a strip scored by torch.max over its patches.  Returns (per-document patch similarities
    (n_b * L,), per-document strip scores (n_b,), pooled questions (B, d))."""
    B = len(patch_embeddings)
    if question_mask is None:
        question_mask = torch.ones(question_embeddings.shape[:2], dtype=torch.int64)
    q = mean_pooling(question_embeddings, question_mask)
    d = question_embeddings.shape[2]
    flat = [p.reshape(-1, d) for p in patch_embeddings]
    sims = score(flat, q)
    strips = []
    for b in range(B):
        n = patch_embeddings[b].shape[0]
        strips.append(sims[b].reshape(n, -1).max(dim=1).values if n else torch.empty(0))
    return sims, strips, q


# --------------------------------------------------------------------------------------
# f2: the generator's input embeddings   reference: src/_modules.py:70-86, src/VT5.py:194-204
# --------------------------------------------------------------------------------------
def spatial_embeddings(bbox: torch.Tensor, x_emb: torch.Tensor, y_emb: torch.Tensor, ln_weight: torch.Tensor,
                       ln_bias: torch.Tensor, eps: float, lin_weight: torch.Tensor, lin_bias: torch.Tensor) -> torch.Tensor:
    """SpatialEmbeddings.forward in eval mode (dropout = identity): the same torch operators in the same order, in the
    dtype of the weights (float32 = the reference; pass float64 copies for the yardstick)."""
    left = F.embedding(bbox[:, :, 0], x_emb)                             # :71
    upper = F.embedding(bbox[:, :, 1], y_emb)                            # :72
    right = F.embedding(bbox[:, :, 2], x_emb)                            # :73
    lower = F.embedding(bbox[:, :, 3], y_emb)                            # :74
    emb = left + upper + right + lower                                   # :76-81
    emb = F.layer_norm(emb, (emb.shape[-1],), ln_weight, ln_bias, eps)   # :83 (BertLayerNorm = nn.LayerNorm)
    return F.linear(emb, lin_weight, lin_bias)                           # :85 (MLP with one layer = nn.Linear)


def vt5_input_embeds(input_ids: torch.Tensor, bbox: torch.Tensor, shared: torch.Tensor, spatial: torch.Tensor,
                     layout_labels: torch.Tensor = None, layout_emb: torch.Tensor = None, layout_scale: float = 1.0):
    """The sum of VT5.prepare_inputs_for_vqa (src/VT5.py:194-204); `spatial` = spatial_embeddings(bbox, ...)."""
    out = F.embedding(input_ids, shared) + spatial                       # :195, :202
    if layout_labels is not None:
        out = out + F.embedding(layout_labels, layout_emb) * layout_scale    # :198, :204
    return out


def late_interaction(query: torch.Tensor, patches: torch.Tensor) -> torch.Tensor:
    qn = F.normalize(query, p=2, dim=-1)                                # :445
    pn = F.normalize(patches, p=2, dim=-1)                              # :446
    S = torch.bmm(qn.expand(pn.size(0), -1, -1), pn.transpose(1, 2))    # :448-451
    return S.max(dim=-1).values.sum(dim=-1)                             # :454-457


def late_interaction_f64(query: torch.Tensor, patches: torch.Tensor) -> torch.Tensor:
    """Same math in float64 -- the yardstick for the fp32 GPU kernel's summation-order error."""
    return late_interaction(query.double(), patches.double())


def visual_scores(patch_embeddings, question_embeddings):
    return [late_interaction(question_embeddings[b].unsqueeze(0), patch_embeddings[b])
            for b in range(len(patch_embeddings))]


# --------------------------------------------------------------------------------------
# a10: visual top-k decode  reference: src/_modules.py:2207-2282, 2284-2384, 2386-2450
# --------------------------------------------------------------------------------------
def surrounding_cells(row: int, col: int, n_rows: int, n_cols: int, include_surroundings):
    cells = set()
    if isinstance(include_surroundings, tuple) and len(include_surroundings) == 2:   # :2237-2244
        rx, ry = include_surroundings
        for r in range(row - ry, row + ry + 1):
            for c in range(col - rx, col + rx + 1):
                cells.add((r, c))
    else:                                                                            # :2246-2274
        level, phase = include_surroundings // 3, include_surroundings % 3
        for r in range(row - level, row + level + 1):
            for c in range(col - level, col + level + 1):
                cells.add((r, c))
        if phase > 0:
            for r in range(row - level, row + level + 1):
                cells.add((r, col - level - 1))
                cells.add((r, col + level + 1))
        if phase > 1:
            for c in range(col - level, col + level + 1):
                cells.add((row - level - 1, c))
                cells.add((row + level + 1, c))
    return {(r, c) for r, c in cells if 0 <= r < n_rows and 0 <= c < n_cols}


def rectangles_overlap(a, b) -> bool:                                                 # src/utils.py:460-463
    return a[0] < b[2] and a[2] > b[0] and a[1] < b[3] and a[3] > b[1]


def merged_rectangles(cells, matrix_shapes, patches_xyxy):
    """cells: iterable of (group, row, col).  Returns {group: sorted list of merged rectangles}
    (bounding box per connected component of the strict-overlap graph; :2331-2382)."""
    by_group = {}
    for g, r, c in cells:
        n_rows, n_cols = matrix_shapes[g]
        if 0 <= r < n_rows and 0 <= c < n_cols:
            by_group.setdefault(g, []).append(list(patches_xyxy[g][r]))
    out = {}
    for g, rects in by_group.items():
        parent = list(range(len(rects)))

        def find(i):
            while parent[i] != i:
                parent[i] = parent[parent[i]]
                i = parent[i]
            return i
        for i in range(len(rects)):
            for j in range(i + 1, len(rects)):
                if rectangles_overlap(rects[i], rects[j]):
                    parent[find(i)] = find(j)
        comps = {}
        for i, rc in enumerate(rects):
            comps.setdefault(find(i), []).append(rc)
        out[g] = sorted([min(x[0] for x in comp), min(x[1] for x in comp),
                         max(x[2] for x in comp), max(x[3] for x in comp)] for comp in comps.values())
    return out


def visual_decode(topk_indices, patches_flatten_indices, matrix_shapes, patches_xyxy,
                  include_surroundings=0, mode: str = "horizontal"):
    """Per document: hit strip index -> (group,row,0) -> neighbourhood -> merged crop rectangles.
    Returns (rects_per_doc: {group: [rect...]}, groups_per_doc: sorted list).  The reference returns
    PIL crops and page ids in Python-set iteration order (:2428, :2445); compare as multisets."""
    rects_all, groups_all = [], []
    for b, hits in enumerate(topk_indices):
        flat = np.asarray(patches_flatten_indices[b])
        if len(flat) == 0:                                                            # :2403-2406
            rects_all.append({})
            groups_all.append([])
            continue
        cells = set()
        for idx in hits:
            idx = int(idx)
            group = int(flat[idx])                                                    # :2411
            row = idx - int(np.count_nonzero(flat < group))                           # :2412
            if mode == "square":
                raise NotImplementedError()                                           # :2413-2414
            n_rows, n_cols = matrix_shapes[b][group]
            for r, c in surrounding_cells(row, 0, n_rows, n_cols, include_surroundings):
                cells.add((group, r, c))
        rects_all.append(merged_rectangles(cells, matrix_shapes[b], patches_xyxy[b]))
        groups_all.append(sorted({g for g, _, _ in cells}))
    return rects_all, groups_all


# --------------------------------------------------------------------------------------
# a12: generator-input assembly   reference: src/utils.py:233-253 (flatten),
#                                  src/VT5.py:141-192 (prepare_inputs_for_vqa, ids/boxes/mask part)
# --------------------------------------------------------------------------------------
def flatten(lst, add_sep_token: Optional[str] = None):
    if add_sep_token is None:
        return [item for sub in lst for item in sub]
    flat = []
    for i, sub in enumerate(lst):
        if len(sub) == 0:
            continue
        if i > 0:
            if isinstance(sub[0], str):
                flat.append(add_sep_token)
            elif isinstance(sub[0], list):
                flat.append([0, 0, 0, 0])
            elif isinstance(sub[0], int):
                flat.append(0)
        flat.extend(sub)
    return flat


def vt5_pack(prompt_token_ids: Sequence[Sequence[int]], words, boxes, word_tokens,
             layout_labels=None, max_source_length: int = 512, eos_id: int = 1, pad_id: int = 0):
    """Packed generator tensors from already-flattened per-document word/box lists.
    `prompt_token_ids[b]` are the prompt ids WITHOUT the trailing EOS (src/VT5.py:147-148);
    `word_tokens(word)` returns the word's ids without EOS (:160)."""
    B = len(words)
    ids_all, boxes_all, lab_all = [], [], []
    longest = 0
    for b in range(B):
        ids = list(prompt_token_ids[b])
        bxs = [[0, 0, 1000, 1000]] * len(ids)                                         # :133, :151
        labs = [4] * len(ids)                                                         # :136, :149
        for i, word in enumerate(words[b]):
            toks = word_tokens(word)
            ids.extend(toks)
            bxs.extend((np.array([boxes[b][i]] * len(toks)) * 1000).tolist())         # :162 (float64)
            if layout_labels is not None:
                labs.extend([layout_labels[b][i]] * len(toks))
        ids_all.append(ids[:max_source_length - 1] + [eos_id])                        # :166
        if len(bxs[:max_source_length - 1]):
            boxes_all.append(np.concatenate([np.array(bxs[:max_source_length - 1], dtype=np.float64),
                                             np.zeros((1, 4))]))                      # :167
        else:
            boxes_all.append(np.zeros((1, 4)))
        lab_all.append(labs[:max_source_length - 1] + [4])                            # :169
        longest = min(max(longest, len(ids) + 1), max_source_length)                  # :170
    t_ids = torch.full([B, longest], pad_id, dtype=torch.long)                        # :173
    t_boxes = torch.zeros([B, longest, 4], dtype=torch.long)                          # :174
    t_labs = torch.full([B, longest], 4, dtype=torch.long)                            # :176
    t_mask = torch.zeros([B, longest], dtype=torch.long)                              # :177
    for b in range(B):
        n = len(ids_all[b])
        t_ids[b, :n] = torch.LongTensor(ids_all[b])
        t_boxes[b, :n] = torch.from_numpy(boxes_all[b][:n])                           # float64 -> int64 truncates
        if layout_labels is not None:
            t_labs[b, :n] = torch.LongTensor(lab_all[b])
        t_mask[b, :n] = 1
    return t_ids, t_boxes, t_mask, (t_labs if layout_labels is not None else None)


# --------------------------------------------------------------------------------------
# corpus mode (C5): sharded top-k + merge   (no reference counterpart: north_star config 5;
#                   the scoring formula is a4's, applied to Q questions x N chunks)
# --------------------------------------------------------------------------------------
def corpus_scores(E: torch.Tensor, Q: torch.Tensor) -> torch.Tensor:
    """(Q, N) cosine with a4's eps convention, in float32 math on float32 copies of E, Q."""
    E32, Q32 = E.float(), Q.float()
    return (Q32 @ E32.T) / (Q32.norm(dim=-1)[:, None] * E32.norm(dim=-1)[None, :] + 1e-8)


def merge_topk(cand_scores: np.ndarray, cand_idx: np.ndarray, k: int):
    """cand_*: (Q, m) candidates (idx < 0 = empty slot) -> (Q, k) by (score desc, idx asc)."""
    Qn = cand_scores.shape[0]
    out_s = np.full((Qn, k), -np.inf, dtype=np.float32)
    out_i = np.full((Qn, k), -1, dtype=np.int64)
    for q in range(Qn):
        valid = cand_idx[q] >= 0
        s, i = cand_scores[q][valid], cand_idx[q][valid].astype(np.int64)
        key = (order_key(s).astype(np.uint64) << np.uint64(32)) | (np.uint64(0xFFFFFFFF) - i.astype(np.uint64))
        order = np.argsort(key, kind="stable")[::-1][:k]
        out_s[q, :len(order)] = s[order]
        out_i[q, :len(order)] = i[order]
    return out_s, out_i


# --------------------------------------------------------------------------------------
# f1: retrieved patches -> the generator's visual input
#   crop rectangles (a9) -> concatenate_patches(mode="grid")   reference: src/utils.py:180-231 (used at
#   src/RAGVT5.py:378 for page_retrieval == "concat", every shipped config) -> the HF feature extractor's
#   resize to 224 x 224 (src/_modules.py:133), which is PIL.Image.resize (third party: Pillow, pinned by the
#   reference's requirements.txt; algorithm restated below from Pillow's src/libImaging/Resample.c, 8 bits
#   per channel path, and checked bit for bit against the installed Pillow in tests/test_oracle_golden.py).
# --------------------------------------------------------------------------------------
def grid_layout(sizes: Sequence[Tuple[int, int]]):
    """Placement of patches of the given (width, height) by concatenate_patches(mode="grid") (src/utils.py:189-231,
    compute_grid :180-187).  Returns (grid_w, grid_h, [(x, y), ...]); no patches -> the 5 x 5 blank image (:193-195)."""
    if not sizes:
        return 5, 5, []
    total_area = sum(w * h for w, h in sizes)                    # :183
    grid_w = max(w for w, _ in sizes)                            # :185
    grid_h = int(total_area / grid_w)                            # :186  (true division, truncation)
    pos, x_off, y_off, row_h = [], 0, 0, 0
    for w, h in sizes:                                           # :223-230, original patch order
        if x_off + w > grid_w:
            x_off = 0
            y_off += row_h
            row_h = 0
        pos.append((x_off, y_off))
        x_off += w
        row_h = max(row_h, h)
    return grid_w, grid_h, pos


def concat_grid(pages: Sequence[np.ndarray], rects: Sequence[Sequence[int]], page_of: Sequence[int]) -> np.ndarray:
    """uint8 (grid_h, grid_w, 3) image = concatenate_patches([page.crop(rect) ...], mode="grid"): crops that reach
    outside the page are black there (PIL crop), pastes are clipped to the canvas (PIL paste)."""
    sizes = [(r[2] - r[0], r[3] - r[1]) for r in rects]
    gw, gh, pos = grid_layout(sizes)
    canvas = np.zeros((gh, gw, 3), dtype=np.uint8)
    for (x0, y0, x1, y1), p, (dx, dy) in zip(rects, page_of, pos):
        page = pages[p]
        H, W = page.shape[:2]
        w, h = x1 - x0, y1 - y0
        patch = np.zeros((max(h, 0), max(w, 0), 3), dtype=np.uint8)
        sx0, sy0, sx1, sy1 = max(x0, 0), max(y0, 0), min(x1, W), min(y1, H)
        if sx1 > sx0 and sy1 > sy0:
            patch[sy0 - y0:sy1 - y0, sx0 - x0:sx1 - x0] = page[sy0:sy1, sx0:sx1]
        cw, ch = min(w, gw - dx), min(h, gh - dy)                 # paste clips at the canvas border
        if cw > 0 and ch > 0:
            canvas[dy:dy + ch, dx:dx + cw] = patch[:ch, :cw]
    return canvas


PIL_BILINEAR, PIL_BICUBIC = 2, 3            # PIL.Image.Resampling values
_PRECISION_BITS = 32 - 8 - 2


def _pil_filter(kind: int, x: float) -> float:
    if x < 0.0:
        x = -x
    if kind == PIL_BILINEAR:
        return 1.0 - x if x < 1.0 else 0.0
    a = -0.5
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def pil_coeffs(in_size: int, out_size: int, kind: int):
    """precompute_coeffs + normalize_coeffs_8bpc of Pillow's Resample.c for the full-image box:
    returns (bounds (out, 2) int32 [first, count], kk (out, ksize) int32 fixed-point weights)."""
    support_f = 1.0 if kind == PIL_BILINEAR else 2.0
    scale = float(np.float32(in_size) - np.float32(0.0)) / out_size
    filterscale = max(scale, 1.0)
    support = support_f * filterscale
    ksize = int(np.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = 0.0 + (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        xmin = max(xmin, 0)
        xmax = int(center + support + 0.5)
        xmax = min(xmax, in_size) - xmin
        w = [_pil_filter(kind, (x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        for x in range(xmax):
            v = w[x] / ww if ww != 0.0 else w[x]
            kk[xx, x] = int(-0.5 + v * (1 << _PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << _PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return bounds, kk


def _pil_pass(src: np.ndarray, bounds: np.ndarray, kk: np.ndarray) -> np.ndarray:
    """One 8-bit resampling pass along axis 1 of src (rows, in, 3) -> (rows, out, 3): 32-bit accumulation from
    1 << (PRECISION_BITS - 1), arithmetic shift, clip to 0..255 (ImagingResampleHorizontal_8bpc)."""
    out = np.empty((src.shape[0], bounds.shape[0], 3), dtype=np.uint8)
    s = src.astype(np.int64)
    for xx in range(bounds.shape[0]):
        x0, n = int(bounds[xx, 0]), int(bounds[xx, 1])
        acc = (s[:, x0:x0 + n, :] * kk[xx, :n].astype(np.int64)[None, :, None]).sum(axis=1) + (1 << (_PRECISION_BITS - 1))
        out[:, xx, :] = np.clip(acc >> _PRECISION_BITS, 0, 255).astype(np.uint8)
    return out


def pil_resize_u8(img: np.ndarray, out_w: int, out_h: int, kind: int = PIL_BICUBIC) -> np.ndarray:
    """PIL.Image.resize((out_w, out_h), resample=kind) of an RGB uint8 image (reducing_gap=None): horizontal pass
    over the rows the vertical pass needs, rounded to uint8, then the vertical pass (ImagingResample)."""
    in_h, in_w = img.shape[:2]
    if (in_w, in_h) == (out_w, out_h):
        return img.copy()
    bh, kh = pil_coeffs(in_w, out_w, kind)
    bv, kv = pil_coeffs(in_h, out_h, kind)
    cur = img
    if out_w != in_w:
        first = int(bv[0, 0])
        last = int(bv[-1, 0] + bv[-1, 1])
        cur = _pil_pass(img[first:last], bh, kh)
        bv = bv.copy()
        bv[:, 0] -= first
    if out_h != in_h:
        cur = _pil_pass(cur.transpose(1, 0, 2), bv, kv).transpose(1, 0, 2)
    return np.ascontiguousarray(cur)


def visual_input(pages, rects, page_of, out_size: int = 224, kind: int = PIL_BICUBIC,
                 mean=(0.5, 0.5, 0.5), std=(0.5, 0.5, 0.5)):
    """(uint8 (S, S, 3) resized grid image, float32 (3, S, S) pixel_values = (u8 / 255 - mean) / std)."""
    u8 = pil_resize_u8(concat_grid(pages, rects, page_of), out_size, out_size, kind)
    px = (u8.astype(np.float32) * np.float32(1.0 / 255.0) - np.asarray(mean, np.float32)) / np.asarray(std, np.float32)
    return u8, np.ascontiguousarray(px.transpose(2, 0, 1))


# --------------------------------------------------------------------------------------
# a12, Pix2Struct half: retrieved image crops -> flattened patches for the Pix2Struct generator
#   reference: src/custom_pix2struct_processor.py:33-132 (extract_flattened_patches_single,
#   extract_multi_image_flattened_patches), :175-196 (per-image normalize), :225 (attention mask);
#   called from src/RAGPix2Struct.py:221.  Third-party pieces: torch.nn.functional.interpolate(bilinear,
#   antialias=True) (called directly below) and transformers==4.49.0 torch_extract_patches (unfold + permute,
#   restated in _extract_patches; the installed transformers 5.x has the same algorithm with a batch dimension).
#   render_header (:214, text drawn onto the first image) is CPU text rendering and out of scope: callers pass the
#   images as they are after that step.
# --------------------------------------------------------------------------------------
def pix2struct_normalize(image: np.ndarray) -> np.ndarray:
    """CustomPix2StructImageProcessor.normalize (:175-196): whole-image mean / std, std floored at 1/sqrt(#elements)."""
    import math
    if image.dtype == np.uint8:
        image = image.astype(np.float32)
    mean = np.mean(image)
    std = np.std(image)
    adjusted = max(std, 1.0 / math.sqrt(np.prod(image.shape)))
    return ((image - mean) / adjusted).astype(np.float32)          # transformers.image_transforms.normalize


def _extract_patches(image: torch.Tensor, ph: int, pw: int) -> torch.Tensor:
    """transformers 4.49 torch_extract_patches: (C, H, W) -> (1, H/ph, W/pw, ph*pw*C), pixel-major / channel-minor."""
    x = image.unsqueeze(0)
    p = F.unfold(x, (ph, pw), stride=(ph, pw))
    p = p.reshape(x.size(0), x.size(1), ph, pw, -1)
    p = p.permute(0, 4, 2, 3, 1).reshape(x.size(2) // ph, x.size(3) // pw, x.size(1) * ph * pw)
    return p.unsqueeze(0)


def pix2struct_patches_single(image: np.ndarray, max_patches: int, ph: int = 16, pw: int = 16, row_offset: int = 0):
    """extract_flattened_patches_single(..., pad=False) for an (H, W, C) float image (:33-95)."""
    import math
    img = torch.from_numpy(np.ascontiguousarray(image.transpose(2, 0, 1)))                 # channels first (:45)
    H, W = img.shape[1], img.shape[2]
    scale = math.sqrt(max_patches * (ph / H) * (pw / W))                                    # :52
    rows = max(min(math.floor(scale * H / ph), max_patches), 1)                             # :53
    cols = max(min(math.floor(scale * W / pw), max_patches), 1)                             # :54
    rh, rw = max(rows * ph, 1), max(cols * pw, 1)
    img = F.interpolate(img.unsqueeze(0), size=(rh, rw), mode="bilinear", align_corners=False, antialias=True).squeeze(0)
    patches = _extract_patches(img, ph, pw)
    r, c, depth = patches.shape[1], patches.shape[2], patches.shape[3]
    patches = patches.reshape(r * c, depth)
    row_ids = torch.arange(r).reshape(r, 1).repeat(1, c).reshape(r * c, 1) + 1 + row_offset   # :79-81
    col_ids = torch.arange(c).reshape(1, c).repeat(r, 1).reshape(r * c, 1) + 1
    result = torch.cat([row_ids.to(torch.float32), col_ids.to(torch.float32), patches], dim=-1)
    return result[:max_patches].numpy(), int(row_ids.max().item())                           # :93-95 (pad=False)


def pix2struct_patches(images: Sequence[np.ndarray], max_total_patches: int = 2048, ph: int = 16, pw: int = 16,
                       normalize: bool = True):
    """extract_multi_image_flattened_patches (:97-132) after the per-image normalize (:220), and the attention mask
    of preprocess (:225).  images: (H, W, 3) uint8 / float arrays.  Returns ((max_total, 2 + ph*pw*3) f32, (max_total,) f32)."""
    if len(images) == 0:
        raise ValueError("No images provided.")                                              # :109
    per = max_total_patches // len(images)                                                   # :110
    out, row_offset = [], 0
    for img in images:
        x = pix2struct_normalize(img) if normalize else np.asarray(img, dtype=np.float32)
        p, row_offset = pix2struct_patches_single(x, per, ph, pw, row_offset)
        out.append(p)
    cat = np.concatenate(out, axis=0)
    if cat.shape[0] < max_total_patches:
        cat = np.concatenate([cat, np.zeros((max_total_patches - cat.shape[0], cat.shape[1]), dtype=cat.dtype)], axis=0)
    else:
        cat = cat[:max_total_patches]
    return cat, (cat.sum(axis=-1) != 0).astype(np.float32)


# --------------------------------------------------------------------------------------
# f3: what consumes the top-k -- reranker post-processing and the page vote
#     reference: src/_modules.py:1562-1610 (Reranker.rerank / batch_rerank), src/RAGVT5.py:455-477 (majorpage /
#     weightmajorpage).  The cross-encoder itself is a model and out of scope: its scores are the input here.
# --------------------------------------------------------------------------------------
def rerank_order(scores, filter_thresh: float = 0.4, max_chunk_num: int = 5, min_chunk_num: int = 1) -> List[int]:
    """The index list Reranker.rerank applies to the candidates and to every extra argument (:1579-1595).
    `np.argsort(scores)[::-1]`: the order of EQUAL scores is unspecified in the reference -- numpy's default sort is
    an insertion sort (stable) up to 16 elements on its scalar path, so ties come out higher index first after the
    reversal, but the AVX-512 / AVX2 argsort numpy dispatches to on recent CPUs (>= 1.25 / 2.0) orders ties
    arbitrarily (seen in tests/golden/postproc.json, made with numpy 2.3).  Stated rule here and in the CUDA kernel:
    ties -> HIGHER index first (kind="stable" reversed); the comparator accepts the reference's order modulo ties.
    NaN sorts last ascending = first after the reversal, never passes the threshold, and can only return through
    the `min_chunk_num` fallback.  The threshold comparison is numpy-scalar >= Python float: float64 under the
    reference's pinned numpy 1.26.4 (legacy promotion); for the shipped 0.4 a float32 comparison agrees
    (float32(0.4) > 0.4 and no float32 lies between)."""
    scores = np.asarray(scores)
    sorted_indices = np.argsort(scores, kind="stable")[::-1]                                   # :1582
    thresh = float(filter_thresh)
    filtered = [int(i) for i in sorted_indices if float(scores[i]) >= thresh]                  # :1585
    if len(filtered) > max_chunk_num:                                                          # :1586-1587
        filtered = filtered[:max_chunk_num]
    elif len(filtered) < min_chunk_num:                                                        # :1588-1589
        filtered = [int(i) for i in sorted_indices[:min_chunk_num]]
    return filtered


def rerank(scores, candidates, *args, filter_thresh: float = 0.4, max_chunk_num: int = 5, min_chunk_num: int = 1):
    """Reranker.rerank after the cross-encoder call (:1592-1595): candidates and every argument permuted alike."""
    order = rerank_order(scores, filter_thresh, max_chunk_num, min_chunk_num)
    return ([candidates[i] for i in order], *[[arg[i] for i in order] for arg in args])


def int_set_order(values: Sequence[int]) -> List[int]:
    """Iteration order of `set(values)` for non-negative ints, i.e. of `list(set(page_indices_b))`
    (src/RAGVT5.py:466) -- the order `max(page_weights, key=...)` breaks ties in (:474).  CPython's
    Objects/setobject.c (3.7 .. 3.12, same algorithm): open addressing, hash(i) = i, 9 linear probes then the
    perturbed jump, table grown to 4 x used when fill * 5 >= mask * 3, entries re-inserted in slot order."""
    LINEAR_PROBES, PERTURB_SHIFT = 9, 5

    def insert(table, mask, v):
        perturb = v
        i = v & mask
        while True:
            probes = LINEAR_PROBES if i + LINEAR_PROBES <= mask else 0
            for j in range(i, i + probes + 1):
                if table[j] is None:
                    table[j] = v
                    return True
                if table[j] == v:
                    return False
            perturb >>= PERTURB_SHIFT
            i = (i * 5 + 1 + perturb) & mask

    mask, table, used = 7, [None] * 8, 0
    for v in values:
        v = int(v)
        if v < 0:
            raise ValueError("page indices are non-negative")
        if insert(table, mask, v):
            used += 1
            if used * 5 >= mask * 3:
                size = 8
                while size <= used * 4:
                    size <<= 1
                old, table, mask = table, [None] * size, size - 1
                for e in old:
                    if e is not None:
                        insert(table, mask, e)
    return [e for e in table if e is not None]


def page_vote(page_indices_b: Sequence[int], similarities_b: Optional[np.ndarray], n_chunks: int, weighted: bool,
              legacy_promotion: bool = True) -> int:
    """major_page_indices[b] of RAGVT5.forward (src/RAGVT5.py:455-475) for one document.
    majorpage: weights = ones(len(similarities[b])) / n  (float64).  weightmajorpage: weights = similarities[b]
    (float32, ALL n_b chunks in chunk order) / sum(w); `zip(page_indices_b, weights_b)` then pairs hit j with the
    weight of CHUNK j (not of the hit) -- reproduced as written.  `sum(w)` and `page_weights[page] += weight` start
    from the Python int 0: with the reference's pinned numpy 1.26.4 int + float32 promotes to float64 (legacy_promotion,
    the default); numpy >= 2 (NEP 50) stays in float32.  The division is float32 in both (array / scalar)."""
    n_hits = len(page_indices_b)
    if weighted:
        w32 = np.asarray(similarities_b, dtype=np.float32)
        if legacy_promotion:
            total = np.float64(0.0)
            for x in w32:
                total = total + np.float64(x)
            weights = w32 / np.float32(total)
            acc_t = np.float64
        else:
            total = np.float32(0.0)
            for x in w32:
                total = np.float32(total + x)
            weights = w32 / total
            acc_t = np.float32
    else:
        weights = np.ones(n_chunks) / float(n_chunks) if n_chunks else np.ones(0)
        acc_t = np.float64
    order = int_set_order(page_indices_b)
    acc = {p: acc_t(0) for p in order}
    for page, weight in zip(page_indices_b, weights[:n_hits]):
        acc[page] = acc_t(acc[page] + acc_t(weight))
    if not acc:
        return 0                                                                               # :471-473
    best = order[0]
    for p in order[1:]:
        if acc[p] > acc[best]:                                                                 # max(): first maximum
            best = p
    return best


# --------------------------------------------------------------------------------------
# f4: Chunker -- words -> layout boxes -> chunks
#     reference: src/utils.py:328-341 (containment_ratio), src/_modules.py:872-1100 (Chunker.get_chunks)
# --------------------------------------------------------------------------------------
def containment_ratio(small_box, large_box) -> float:                                         # src/utils.py:328-341
    x1 = max(small_box[0], large_box[0]); y1 = max(small_box[1], large_box[1])
    x2 = min(small_box[2], large_box[2]); y2 = min(small_box[3], large_box[3])
    inter_area = max(0, x2 - x1) * max(0, y2 - y1)
    small_area = (small_box[2] - small_box[0]) * (small_box[3] - small_box[1])
    return inter_area / small_area if small_area > 0 else 0


class ChunkStats:
    """The four counters Chunker keeps (src/_modules.py:855-866) as plain dicts; zero entries stay, as with the
    reference's subtract-then-add bookkeeping."""
    def __init__(self):
        self.chunk_size_dist, self.n_chunks_per_page_dist = {}, {}
        self.n_chunks_per_doc_dist, self.n_chunks_per_layout_dist = {}, {}

    @staticmethod
    def add(d, key, value=1):
        d[key] = d.get(key, 0) + value

    def as_dict(self):
        return {k: {str(a): b for a, b in getattr(self, k).items()} for k in
                ("chunk_size_dist", "n_chunks_per_page_dist", "n_chunks_per_doc_dist", "n_chunks_per_layout_dist")}


def make_chunks(words, boxes, tag, words_lst, boxes_lst, tag_lst, chunk_size, overlap, tol, stats: ChunkStats) -> int:
    """The closure of get_chunks (:906-938): windows of chunk_size words every chunk_size - overlap words; a window is
    folded into the previous chunk when the sizes -- by the reference's own arithmetic, prev + len(window) - overlap,
    which can go below the true length -- stay within chunk_size * (1 + tol)."""
    prev, made = 0, 0
    for i in range(0, len(words), chunk_size - overlap):
        cw, cb = words[i:i + chunk_size], boxes[i:i + chunk_size]
        size = len(cw)
        if i > 0 and tag == tag_lst[-1] and prev + (size - overlap) <= chunk_size * (1 + tol):
            size = prev + size - overlap
            words_lst[-1].extend(cw[overlap:]); boxes_lst[-1].extend(cb[overlap:])
            stats.add(stats.chunk_size_dist, prev, -1); stats.add(stats.chunk_size_dist, size)
        else:
            tag_lst.append(tag); words_lst.append(cw); boxes_lst.append(cb)
            stats.add(stats.chunk_size_dist, len(cw))
            made += 1
        prev = size
    return made


def get_chunks(words, boxes, layout_info=None, chunk_size: int = 60, overlap: int = 10, tol: float = 0.2,
               page_retrieval: str = "concat", default_label: int = 1, cluster_layouts: bool = False):
    """Chunker.get_chunks (:872-1100).  Returns the reference's 5-tuple and the counters."""
    from collections import Counter
    stats = ChunkStats()
    bs = len(words)
    lay_boxes = lay_labels = lay_clusters = None
    if layout_info != [[]] and layout_info is not None:                                       # :892-898
        lay_boxes = [[pg["boxes"] for pg in layout_info[b]] for b in range(bs)]
        lay_labels = [[pg["labels"] for pg in layout_info[b]] for b in range(bs)]
        if "clusters" in layout_info[0][0].keys() and cluster_layouts:
            lay_clusters = [[pg["clusters"] for pg in layout_info[b]] for b in range(bs)]
    out_labels, out_pages, out_words, out_boxes, out_word_labels = [], [], [], [], []
    for b in range(bs):
        d_labels, d_pages, d_words, d_boxes, d_word_labels, d_n = [], [], [], [], [], 0
        for p, (page_words, page_boxes) in enumerate(zip(words[b], boxes[b])):
            if not isinstance(page_words, list):                                              # :957-960
                page_boxes = page_boxes.tolist()
            if len(page_boxes) > 0 and not isinstance(page_boxes[0], list):
                page_boxes = [pb.tolist() for pb in page_boxes]
            if page_retrieval == "oracle":                                                    # :962-974
                d_pages.append(p); d_words.append(page_words); d_boxes.append(page_boxes)
                d_labels.append(default_label); d_word_labels.append([default_label] * len(page_words))
                d_n += 1
                stats.add(stats.chunk_size_dist, len(page_words)); stats.add(stats.n_chunks_per_page_dist, 1)
                continue
            if lay_boxes is None or len(lay_boxes[b][p]) == 0:                                # :976-990
                n = make_chunks(page_words, page_boxes, p, d_words, d_boxes, d_pages, chunk_size, overlap, tol, stats)
                d_labels.extend([default_label] * n); d_word_labels.append([default_label] * len(page_words))
                d_n += n
                stats.add(stats.n_chunks_per_page_dist, n)
                continue
            pl_boxes, pl_labels = lay_boxes[b][p], lay_labels[b][p]
            pl_clusters = lay_clusters[b][p].tolist() if lay_clusters else None
            order = sorted(range(len(pl_boxes)), key=lambda j: (pl_boxes[j][0], pl_boxes[j][1]))   # :1006-1018, stable
            pl_boxes = [pl_boxes[j] for j in order]; pl_labels = [pl_labels[j] for j in order]
            if pl_clusters:
                pl_clusters = [pl_clusters[j] for j in order]
            word_labels = [default_label] * len(page_words)
            inside_w, inside_b = [], []
            for lbox, llabel in zip(pl_boxes, pl_labels):                                     # :1023-1033
                ws, bx = [], []
                for i, (word, box) in enumerate(zip(page_words, page_boxes)):
                    if containment_ratio(box, lbox) > 0.5:
                        ws.append(word); bx.append(box); word_labels[i] = llabel
                inside_w.append(ws); inside_b.append(bx)
            group_labels = list(pl_labels)
            if pl_clusters:                                                                   # :1035-1063
                cw, cb, cl, slot = [], [], [], {}
                for ws, bx, llabel, cluster in zip(inside_w, inside_b, pl_labels, pl_clusters):
                    if cluster == -1 or cluster not in slot:
                        if cluster != -1:
                            slot[cluster] = len(cw)
                        cw.append(ws); cb.append(bx); cl.append(Counter([llabel]))
                    else:
                        j = slot[cluster]
                        cw[j].extend(ws); cb[j].extend(bx); cl[j][llabel] += 1
                inside_w, inside_b = cw, cb
                group_labels = [c.most_common(1)[0][0] for c in cl]
            l_words, l_boxes, l_tags, page_n = [], [], [], 0
            for lb, (ws, bx, glabel) in enumerate(zip(inside_w, inside_b, group_labels)):      # :1064-1076
                n = make_chunks(ws, bx, lb, l_words, l_boxes, l_tags, chunk_size, overlap, tol, stats)
                page_n += n
                d_labels.extend([glabel] * n)
                stats.add(stats.n_chunks_per_layout_dist, n)
            d_pages.extend([p] * len(l_words)); d_words.extend(l_words); d_boxes.extend(l_boxes)
            d_word_labels.append(word_labels)
            d_n += page_n
            stats.add(stats.n_chunks_per_page_dist, page_n)
        out_labels.append(d_labels); out_pages.append(d_pages); out_words.append(d_words); out_boxes.append(d_boxes)
        out_word_labels.append(d_word_labels)
        stats.add(stats.n_chunks_per_doc_dist, d_n)
    return (out_words, out_boxes, out_labels, out_pages, out_word_labels), stats


# --------------------------------------------------------------------------------------
# 8f rank 4 (second half): S2Chunker -- layout regions of a page as graph nodes, pairwise weights,
# spectral clustering          reference: src/_modules.py:1669-1962
# Pinned by oracle/make_golden_s2chunker.py -> tests/golden/s2chunker.json (tests/test_s2chunker_oracle.py).
# --------------------------------------------------------------------------------------
def s2_nodes(page_layout_info: dict, page_info: Optional[dict], cluster_mode: str):
    """create_nodes_and_edges (:1687-1753): (nodes, edges, used).  In "spatial+semantic" mode with page_info the
    reference's word loop reuses the node counter `i` (:1724), so global ids start at len(page words) - 1 -- restated
    as written (cluster() then fails in _add_weights_to_graph exactly as the reference does)."""
    boxes, labels = page_layout_info["boxes"], page_layout_info["labels"]
    nodes, used = [], np.zeros(len(boxes), dtype=bool)
    i = 0
    if cluster_mode == "spatial" or page_info is None:
        for l, (box, label) in enumerate(zip(boxes, labels)):
            nodes.append({"global_id": i, "page": 1, "bbox": box, "text": "", "label": label})
            i += 1
            used[l] = True
    else:
        page_words, page_boxes = page_info["ocr_tokens"], page_info["ocr_normalized_boxes"]
        inside = []
        for box in boxes:
            inside.append([w for w, wb in zip(page_words, page_boxes) if containment_ratio(wb, box) > 0.5])   # :1724-1730
        if len(boxes) and len(page_words):
            i = len(page_words) - 1                                                     # the shadowed loop variable
        for l, (box, label) in enumerate(zip(boxes, labels)):
            if not inside[l]:
                continue
            nodes.append({"global_id": i, "page": 1, "bbox": box, "text": " ".join(inside[l]), "label": label})
            i += 1
            used[l] = True
    edges = [(nodes[a]["global_id"], nodes[b]["global_id"]) for a in range(len(nodes)) for b in range(a + 1, len(nodes))]
    return nodes, edges, used


def s2_spatial_weights(boxes) -> np.ndarray:
    """_spatial_weights_calculation (:1755-1773), the reference's own numpy calls per entry."""
    n = len(boxes)
    out = np.zeros((n, n))
    for a in range(n):
        for b in range(n):
            ba, bb = boxes[a], boxes[b]
            ca = np.array([(ba[0] + ba[2]) / 2, (ba[1] + ba[3]) / 2])
            cb = np.array([(bb[0] + bb[2]) / 2, (bb[1] + bb[3]) / 2])
            out[a, b] = 1 / (1 + np.linalg.norm(ca - cb))
    return out


def s2_semantic_weights(embeddings: np.ndarray) -> np.ndarray:
    """_semantic_weights_calculation (:1775-1788): sklearn.metrics.pairwise.cosine_similarity of the node embeddings."""
    from sklearn.metrics.pairwise import cosine_similarity
    return cosine_similarity(embeddings)


def s2_combined_weights(boxes, embeddings: Optional[np.ndarray] = None) -> np.ndarray:
    """_combined_weights (:1790-1802); embeddings None = cluster_mode "spatial"."""
    spatial = s2_spatial_weights(boxes)
    semantic = spatial if embeddings is None else s2_semantic_weights(embeddings)
    return (spatial + semantic) / 2


def s2_best_clusters(weights: np.ndarray, n_nodes: int, min_k: int = 2, max_k: int = 10):
    """_calculate_n_clusters with calculate_n_clusters == "best" (:1815-1849)."""
    from sklearn.cluster import SpectralClustering
    from sklearn.metrics import silhouette_score
    degree = np.sum(weights, axis=1)
    d_inv_sqrt = np.diag(1.0 / (np.sqrt(degree) + 1e-10))
    l_norm = np.eye(weights.shape[0]) - d_inv_sqrt @ weights @ d_inv_sqrt
    _, eigenvectors = np.linalg.eigh(l_norm)
    embedding = eigenvectors[:, :max_k]
    best_k, best_score, best_labels = min_k, -1, np.full(n_nodes, -1)
    for k in range(min_k, min(max_k, n_nodes - 1) + 1):
        labels = SpectralClustering(n_clusters=k, affinity="precomputed").fit_predict(weights)
        score_k = silhouette_score(embedding, labels)
        if score_k > best_score:
            best_score, best_k, best_labels = score_k, k, labels
    return best_k, best_labels


def s2_forward(layout_info: Sequence[dict], pages_info=None, cluster_mode: str = "spatial", embed=None) -> List[np.ndarray]:
    """S2Chunker.forward (:1929-1962) for calculate_n_clusters == "best" (the shipped setting, precompute_layouts.py:130-131)."""
    out = []
    for p, page in enumerate(layout_info):
        if len(page["boxes"]) == 0:
            out.append(np.array([]))
            continue
        page_info = pages_info[p] if pages_info is not None else None
        nodes, edges, used = s2_nodes(page, page_info, cluster_mode)
        if len(nodes) < 2:
            out.append(np.full(len(page["boxes"]), -1))
            continue
        emb = None
        if cluster_mode == "spatial+semantic":
            emb = embed([n["text"] for n in nodes if n.get("text", "").strip()])
        weights = s2_combined_weights([n["bbox"] for n in nodes], emb)
        ids = [n["global_id"] for n in nodes]
        for (u, v) in edges:                                # _add_weights_to_graph (:1810-1813): weights[u, v] by GLOBAL id
            weights[u, v]
        _, labels = s2_best_clusters(weights, len(nodes))
        clusters = [lab for _, lab in sorted(zip(ids, labels), key=lambda item: item[0])]
        complete = np.full(len(used), -1)
        complete[used] = clusters
        out.append(complete)
    return out
