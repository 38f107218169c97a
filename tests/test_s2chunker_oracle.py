"""Oracle restatement of S2Chunker (second half) against the reference's frozen outputs
(tests/golden/s2chunker.json, oracle/make_golden_s2chunker.py), against the live reference when it is present, and the
host half of the product's S2Chunker (everything but the two kernels) against the same files."""
import json
import os
import zlib

import numpy as np
import pytest

from oracle import ref_restated as R
from oracle.ref_import import import_reference, reference_available
from rag_docvqa_b200 import synth


def load_cases(golden_dir):
    with open(os.path.join(golden_dir, "s2chunker.json")) as f:
        return json.load(f)


def from_hex(values, n):
    return np.array([float.fromhex(v) for v in values], dtype=np.float64).reshape(n, n)


def test_nodes_and_weights_match_reference_golden(golden_dir):
    emb = synth.HashEmbedder(384)
    n_matrices = 0
    for rec in load_cases(golden_dir):
        layout_info, pages_info = synth.make_s2_pages(**rec["case"])
        for item in rec["pages"]:
            mode, p = item["mode"], item["page"]
            nodes, edges, used = R.s2_nodes(layout_info[p], pages_info[p] if mode != "spatial" else None, mode)
            assert [n["global_id"] for n in nodes] == item["ids"]
            assert [bool(u) for u in used] == item["used"] and len(edges) == item["n_edges"]
            assert [zlib.crc32(n["text"].encode()) for n in nodes] == item["texts_crc"]
            if "weights" not in item:
                continue
            want = from_hex(item["weights"], len(nodes))
            e = emb.forward([n["text"] for n in nodes]).numpy() if mode != "spatial" else None
            got = R.s2_combined_weights([n["bbox"] for n in nodes], e)
            if mode == "spatial":
                assert np.array_equal(got, want), (rec["case"], p)          # the reference's own numpy calls: bit-exact
            else:
                np.testing.assert_allclose(got, want, rtol=0, atol=1e-6)    # BLAS sgemm inside sklearn
            n_matrices += 1
    assert n_matrices >= 30


def test_spatial_forward_matches_reference_golden(golden_dir):
    for rec in load_cases(golden_dir):
        layout_info, _ = synth.make_s2_pages(**rec["case"])
        np.random.seed(0)
        got = R.s2_forward(layout_info, None, "spatial")
        assert [np.asarray(c).astype(int).tolist() for c in got] == rec["clusters_spatial_best"]


def test_semantic_mode_with_page_words_fails_like_the_reference():
    """Global ids start at len(words) - 1 (:1724), so weights[u, v] (:1812) is out of range for any real page."""
    layout_info, pages_info = synth.make_s2_pages(seed=21, pages=2, max_layouts=6, max_words=80, degenerate=False)
    with pytest.raises(IndexError):
        R.s2_forward(layout_info, pages_info, "spatial+semantic", embed=lambda t: synth.HashEmbedder(32).forward(t).numpy())


@pytest.mark.skipif(not reference_available(), reason="reference tree not present")
def test_oracle_against_live_reference():
    modules, _, _ = import_reference()
    emb = synth.HashEmbedder(64)
    emb.bge_model = type("M", (), {"tokenizer": None})()
    layout_info, pages_info = synth.make_s2_pages(seed=31, pages=7, max_layouts=12, max_words=150)
    for mode in ("spatial", "spatial+semantic"):
        s2 = modules.S2Chunker({"cluster_mode": mode, "calculate_n_clusters": "best"}, embedder=emb)
        for p, page in enumerate(layout_info):
            if not len(page["boxes"]):
                continue
            info = pages_info[p] if mode != "spatial" else None
            nodes, edges, used = s2.create_nodes_and_edges(page, info)
            mine = R.s2_nodes(page, info, mode)
            assert nodes == mine[0] and edges == mine[1] and used.tolist() == mine[2].tolist()
            if nodes:
                e = emb.forward([n["text"] for n in nodes]).numpy() if mode != "spatial" else None
                np.testing.assert_allclose(R.s2_combined_weights([n["bbox"] for n in nodes], e), s2._combined_weights(nodes),
                                           rtol=0, atol=0 if mode == "spatial" else 1e-6)
    s2 = modules.S2Chunker({"cluster_mode": "spatial", "calculate_n_clusters": "best"})
    np.random.seed(5)
    want = s2.forward(layout_info)
    np.random.seed(5)
    got = R.s2_forward(layout_info, None, "spatial")
    assert [np.asarray(c).tolist() for c in got] == [np.asarray(c).tolist() for c in want]


def test_product_host_half_needs_no_gpu(golden_dir, monkeypatch):
    """rag_docvqa_b200.s2chunker.S2Chunker with its weight launch replaced by the oracle's matrices: node building in
    "spatial" mode, graph weights and the sklearn clustering are host code and must reproduce the frozen cluster arrays."""
    from rag_docvqa_b200.s2chunker import S2Chunker
    monkeypatch.setattr(S2Chunker, "weights_batch",
                        lambda self, boxes, emb=None, what=0: [R.s2_combined_weights(b, None) for b in boxes])
    for rec in load_cases(golden_dir):
        layout_info, _ = synth.make_s2_pages(**rec["case"])
        s2 = S2Chunker({"cluster_mode": "spatial", "calculate_n_clusters": "best", "device": "cuda:0"})
        np.random.seed(0)
        got = s2.forward(layout_info)
        assert [np.asarray(c).astype(int).tolist() for c in got] == rec["clusters_spatial_best"]
        for p, page in enumerate(layout_info):
            nodes, edges, used = s2.create_nodes_and_edges(page)
            ref = R.s2_nodes(page, None, "spatial")
            assert nodes == ref[0] and edges == ref[1] and used.tolist() == ref[2].tolist()
    with pytest.raises(ValueError):
        S2Chunker({"cluster_mode": "spatial+semantic", "device": "cuda:0"})


@pytest.mark.skipif(not reference_available(), reason="reference tree not present")
def test_product_heuristic_mode_equals_live_reference(monkeypatch):
    """calculate_n_clusters == "heuristic" (KMeans on the spectral embedding, SpectralClustering, then the split by token
    length): the reference needs `max_token_length` set by hand (:1678 is commented out) and a tokenizer; with both given,
    the product's host half (weights from the oracle) must return the reference's arrays under the same numpy seed."""
    from rag_docvqa_b200.s2chunker import S2Chunker
    modules, _, _ = import_reference()
    monkeypatch.setattr(S2Chunker, "weights_batch",
                        lambda self, boxes, emb=None, what=0: [R.s2_combined_weights(b, None) for b in boxes])

    class Tok:
        @staticmethod
        def tokenize(text):
            return text.split() or ["x"] * 7

    layout_info, _ = synth.make_s2_pages(seed=51, pages=5, max_layouts=10, max_words=60)
    ref = modules.S2Chunker({"cluster_mode": "spatial", "calculate_n_clusters": "heuristic"})
    mine = S2Chunker({"cluster_mode": "spatial", "calculate_n_clusters": "heuristic", "device": "cuda:0"})
    for s2 in (ref, mine):
        s2.tokenizer, s2.max_token_length = Tok(), 20
    np.random.seed(3)
    want = ref.forward(layout_info)
    np.random.seed(3)
    got = mine.forward(layout_info)
    assert [np.asarray(c).tolist() for c in got] == [np.asarray(c).tolist() for c in want]
    # without max_token_length both fail the same way
    del mine.max_token_length
    with pytest.raises(AttributeError):
        mine.forward(layout_info)
