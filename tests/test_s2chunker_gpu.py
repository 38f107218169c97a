"""GPU parity of the S2Chunker drop-in (second half): rdv_s2_weights and the batched node building
against the reference's frozen outputs (tests/golden/s2chunker.json) and the oracle."""
import json
import os
import zlib

import numpy as np
import pytest
import torch

from oracle import ref_restated as R
from rag_docvqa_b200 import _lib, synth

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def load_cases(golden_dir):
    with open(os.path.join(golden_dir, "s2chunker.json")) as f:
        return json.load(f)


def from_hex(values, n):
    return np.array([float.fromhex(v) for v in values], dtype=np.float64).reshape(n, n)


def make(mode, embedder=None):
    from rag_docvqa_b200.s2chunker import S2Chunker
    return S2Chunker({"cluster_mode": mode, "calculate_n_clusters": "best", "device": DEV}, embedder=embedder)


def test_nodes_and_weights_match_reference_golden(golden_dir):
    emb = synth.HashEmbedder(384, device=DEV)
    checked = 0
    for rec in load_cases(golden_dir):
        layout_info, pages_info = synth.make_s2_pages(**rec["case"])
        for mode in ("spatial", "spatial+semantic"):
            s2 = make(mode, emb)
            built = s2._nodes_batch(layout_info, pages_info if mode != "spatial" else None)     # one launch for all pages
            items = [it for it in rec["pages"] if it["mode"] == mode]
            assert [it["page"] for it in items] == [p for p, page in enumerate(layout_info) if len(page["boxes"])]
            todo = []
            for it in items:
                nodes, edges, used = built[it["page"]]
                assert [n["global_id"] for n in nodes] == it["ids"]
                assert [bool(u) for u in used] == it["used"] and len(edges) == it["n_edges"]
                assert [zlib.crc32(n["text"].encode()) for n in nodes] == it["texts_crc"]
                single = s2.create_nodes_and_edges(layout_info[it["page"]], pages_info[it["page"]] if mode != "spatial" else None)
                assert single[0] == nodes and single[1] == edges
                if "weights" in it:
                    todo.append((it, nodes))
            embs = [emb.forward([n["text"] for n in nodes]) for _, nodes in todo] if mode != "spatial" else None
            got = s2.weights_batch([[n["bbox"] for n in nodes] for _, nodes in todo], embs)     # one launch for all matrices
            for (it, nodes), w in zip(todo, got):
                want = from_hex(it["weights"], len(nodes))
                if mode == "spatial":
                    assert np.array_equal(w, want), (rec["case"], it["page"])                   # float64, bit-exact
                    assert np.array_equal(s2._combined_weights(nodes), want)
                    assert np.array_equal(s2._spatial_weights_calculation(nodes), want)
                else:
                    np.testing.assert_allclose(w, want, rtol=0, atol=1e-6)
                checked += 1
    assert checked >= 30


def clusters_from_frozen_weights(rec, layout_info):
    """The oracle's clustering run on THIS host from the reference's frozen weight matrices (same seed, same page order)."""
    frozen = {it["page"]: it for it in rec["pages"] if it["mode"] == "spatial"}
    out = []
    np.random.seed(0)
    for p, page in enumerate(layout_info):
        n = len(page["boxes"])
        if n == 0:
            out.append([])
        elif n < 2:
            out.append([-1] * n)
        else:
            _, labels = R.s2_best_clusters(from_hex(frozen[p]["weights"], n), n)
            out.append(np.asarray(labels).astype(int).tolist())
    return out


def test_spatial_forward_matches_reference_golden(golden_dir):
    for rec in load_cases(golden_dir):
        layout_info, _ = synth.make_s2_pages(**rec["case"])
        s2 = make("spatial")
        np.random.seed(0)
        got = [np.asarray(c).astype(int).tolist() for c in s2.forward(layout_info)]
        if got != rec["clusters_spatial_best"]:
            # The labels come out of sklearn (LAPACK eigh, ARPACK, k-means) on the HOST: another CPU model may round them
            # differently than the build container did when the file was frozen.  The weights are bit-exact (tested
            # above), so the same clustering run here on the frozen matrices is the reference's answer on this host.
            assert got == clusters_from_frozen_weights(rec, layout_info)


def test_weight_terms_vs_oracle_larger():
    """Pages of up to 60 regions, 768-d embeddings incl. a zero row: each term of the matrix on its own."""
    rng = np.random.RandomState(3)
    layout_info, _ = synth.make_s2_pages(seed=41, pages=12, max_layouts=60, max_words=8)
    pages = [p["boxes"] for p in layout_info]
    embs = [torch.from_numpy(rng.randn(len(b), 768).astype(np.float32) + 0.3) for b in pages]
    for e in embs:
        if len(e) > 2:
            e[1] = 0.0
    s2 = make("spatial")
    spatial = s2.weights_batch(pages, None, _lib.S2_SPATIAL)
    semantic = s2.weights_batch(pages, embs, _lib.S2_SEMANTIC)
    combined = s2.weights_batch(pages, embs, _lib.S2_COMBINED)
    for b, e, sp, se, co in zip(pages, embs, spatial, semantic, combined):
        n = len(b)
        assert sp.shape == se.shape == co.shape == (n, n)
        if n == 0:
            continue
        ref_sp = R.s2_spatial_weights(b)
        np.testing.assert_allclose(sp, ref_sp, rtol=3e-16, atol=0)          # <= 1 ulp: numpy's ddot may or may not fuse
        np.testing.assert_allclose(se, R.s2_semantic_weights(e.numpy()), rtol=0, atol=1e-6)
        assert np.array_equal(co, (sp + se) / 2)
        assert np.array_equal(sp, sp.T) and np.all(np.diag(sp) == 1.0)
    with pytest.raises(ValueError):
        s2.weights_batch(pages[:1], [embs[0][:-1]])


def test_semantic_mode_with_page_words_fails_like_the_reference():
    layout_info, pages_info = synth.make_s2_pages(seed=21, pages=2, max_layouts=6, max_words=80, degenerate=False)
    s2 = make("spatial+semantic", synth.HashEmbedder(32, device=DEV))
    with pytest.raises(IndexError):
        s2.forward(layout_info, pages_info)
    # without page words the mode works (nodes = all regions, ids from 0; every text empty -> no embeddings: ValueError
    # from the shape mismatch, as numpy raises in the reference at :1801)
    with pytest.raises(ValueError):
        s2.forward(layout_info, None)
