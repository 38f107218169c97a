"""Oracle restatement of Chunker.get_chunks against the reference's frozen outputs
(tests/golden/chunker.json, oracle/make_golden_chunker.py) and the host-side chunk windows of the product."""
import json
import os
import random

import pytest

from oracle import ref_restated as R
from oracle.make_golden_chunker import crc, inputs_of
from oracle.ref_import import reference_available


def load_cases(golden_dir):
    with open(os.path.join(golden_dir, "chunker.json")) as f:
        return json.load(f)


def run_oracle(case):
    words, boxes, info = inputs_of(case)
    return R.get_chunks(words, boxes, info, chunk_size=case["chunk_size"], overlap=case["overlap"], tol=case["tol"],
                        page_retrieval=case["page_retrieval"], cluster_layouts=case["cluster_layouts"])


def test_get_chunks_matches_reference_golden(golden_dir):
    cases = load_cases(golden_dir)
    assert sum(sum(c["n_chunks"]) for c in cases) > 300
    for rec in cases:
        res, stats = run_oracle(rec["case"])
        res = json.loads(json.dumps(res))
        assert [crc(x) for x in res] == rec["crc"], rec["case"]
        if "outputs" in rec:
            assert res == rec["outputs"]
        got = stats.as_dict()
        for key, want in rec["stats"].items():
            assert got[key] == want, (rec["case"], key)


def test_containment_ratio_conventions():
    assert R.containment_ratio([0, 0, 2, 2], [1, 0, 3, 2]) == 0.5          # exactly half: NOT inside (> 0.5)
    assert R.containment_ratio([0, 0, 0, 2], [0, 0, 3, 2]) == 0            # zero-area word
    assert R.containment_ratio([0.1, 0.1, 0.2, 0.2], [0.5, 0.5, 0.9, 0.9]) == 0
    assert R.containment_ratio([0.1, 0.1, 0.2, 0.2], [0.0, 0.0, 1.0, 1.0]) == 1.0


def test_chunk_ranges_equal_the_window_loop():
    """rag_docvqa_b200.chunker.chunk_ranges (index ranges) against the oracle's list-extension loop."""
    from rag_docvqa_b200.chunker import chunk_ranges
    rnd = random.Random(2)
    for _ in range(3000):
        c = rnd.randint(2, 40)
        o = rnd.randint(0, c - 1)
        tol = rnd.choice([0.0, 0.2, 0.5, 1.0])
        n = rnd.randint(0, 300)
        words = list(range(n))
        wl, bl, tl, stats = [], [], [], R.ChunkStats()
        made = R.make_chunks(words, words, 0, wl, bl, tl, c, o, tol, stats)
        ranges, events = chunk_ranges(n, c, o, tol)
        assert [words[a:b] for a, b in ranges] == wl and made == len(ranges)
        mine = {}
        for size, delta in events:
            mine[size] = mine.get(size, 0) + delta
        assert mine == stats.chunk_size_dist


@pytest.mark.skipif(not reference_available(), reason="reference tree not mounted (GPU box)")
def test_oracle_against_live_reference_random():
    from oracle.ref_import import import_reference
    from oracle.make_golden_chunker import config_of, stats_json
    from rag_docvqa_b200 import synth
    modules, _, _ = import_reference()
    for seed in range(20, 26):
        case = dict(seed=seed, docs=2, max_pages=3, max_words=150, max_layouts=9, clusters=seed % 2 == 0,
                    cluster_layouts=seed % 2 == 0, chunk_size=25, overlap=seed % 7, tol=0.2, page_retrieval="concat")
        words, boxes, info = inputs_of(case)
        ch = modules.Chunker(config_of(case))
        ref = ch.get_chunks(words, boxes, info, question_id=["q"] * len(words))
        got, stats = run_oracle(case)
        assert json.loads(json.dumps(got)) == json.loads(json.dumps(ref))
        want = stats_json(ch.stats)
        assert {k: v for k, v in stats.as_dict().items() if k in want} == want


@pytest.mark.parametrize("page_retrieval", ["concat", "oracle"])
def test_product_chunker_host_paths_need_no_gpu(page_retrieval):
    """Without layout boxes (or with page_retrieval == "oracle") Chunker.get_chunks launches nothing: the host logic
    (box normalisation, chunk windows as index ranges, counters) against the oracle's list-extension loop."""
    from rag_docvqa_b200 import synth
    from rag_docvqa_b200.chunker import Chunker
    words, boxes, info = synth.make_chunker_batch(8, 3, 4, 200, 8, numpy_pages=True)
    cfg = {"compute_stats": True, "compute_stats_examples": True, "n_stats_examples": 2, "layout_model_weights": None,
           "device": "cuda:0", "page_retrieval": page_retrieval, "chunk_size": 25, "overlap": 5}
    ch = Chunker(cfg)
    layout = info if page_retrieval == "oracle" else [[]]
    got = ch.get_chunks(words, boxes, layout, question_id=["a", "b", "c"])
    want, stats = R.get_chunks(words, boxes, layout, chunk_size=25, overlap=5, page_retrieval=page_retrieval)
    as_lists = lambda x: json.loads(json.dumps(x, default=lambda o: o.tolist()))     # a page without words stays an empty ndarray
    assert as_lists(got) == as_lists(want)
    assert dict(ch.stats["chunk_size_dist"]) == stats.chunk_size_dist
    assert dict(ch.stats["n_chunks_per_page_dist"]) == stats.n_chunks_per_page_dist
    assert dict(ch.stats["n_chunks_per_doc_dist"]) == stats.n_chunks_per_doc_dist
    assert all(len(v) <= 2 for v in ch.stats_examples["chunk_size_dist"].values())


def emulated_assign(self, page_boxes, layout_boxes, layout_labels, default_label=-1):
    """Stand-in for Chunker.assign_words_to_layouts (one rdv_layout_assign launch on the GPU): the same two outputs from
    the oracle's containment_ratio in Python floats -- the decisions the kernel is tested to reproduce bit for bit
    (tests/test_chunker_gpu.py::test_layout_assign_bit_exact_on_arbitrary_floats)."""
    import numpy as np
    inside, labels = [], []
    for pb, lb, ll in zip(page_boxes, layout_boxes, layout_labels):
        assert pb.dtype == np.float64 and lb.dtype == np.float64 and pb.shape[1:] == (4,) and lb.shape[1:] == (4,)
        ins = np.zeros((len(lb), len(pb)), dtype=bool)
        lab = np.full(len(pb), default_label, dtype=np.int32)
        for g, box in enumerate(lb.tolist()):
            for w, wb in enumerate(pb.tolist()):
                if R.containment_ratio(wb, box) > 0.5:
                    ins[g, w] = True
                    lab[w] = ll[g]
        inside.append(ins)
        labels.append(lab)
    return inside, labels


def test_product_get_chunks_host_half_matches_reference_golden(golden_dir, monkeypatch):
    """Chunker.get_chunks with layout boxes, everything but the kernel: box normalisation, the stable (xmin, ymin) order,
    cluster grouping, chunk windows, word labels and counters against the reference's frozen outputs, on CPU."""
    from oracle.make_golden_chunker import config_of
    from rag_docvqa_b200.chunker import Chunker
    monkeypatch.setattr(Chunker, "assign_words_to_layouts", emulated_assign)
    for rec in load_cases(golden_dir):
        case = rec["case"]
        words, boxes, info = inputs_of(case)
        ch = Chunker({**config_of(case), "device": "cuda:0"})
        res = ch.get_chunks(words, boxes, info, question_id=["q%d" % b for b in range(len(words))])
        res = json.loads(json.dumps(res))
        assert [crc(x) for x in res] == rec["crc"], case
        if "outputs" in rec:
            assert res == rec["outputs"]
        got = {k: {str(a): int(b) for a, b in v.items()} for k, v in ch.stats.items()}
        assert got == rec["stats"], case
