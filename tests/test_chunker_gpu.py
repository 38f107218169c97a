"""GPU parity of the Chunker drop-in: rdv_layout_assign against Python-float containment
ratios, Chunker.get_chunks against the reference's frozen outputs (tests/golden/chunker.json) and the oracle."""
import json
import os

import numpy as np
import pytest

from oracle import ref_restated as R
from oracle.make_golden_chunker import config_of, crc, inputs_of

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def make_chunker(case, **extra):
    from rag_docvqa_b200.chunker import Chunker
    return Chunker({**config_of(case), "device": DEV, **extra})


def test_get_chunks_matches_reference_golden(golden_dir):
    with open(os.path.join(golden_dir, "chunker.json")) as f:
        cases = json.load(f)
    for rec in cases:
        case = rec["case"]
        words, boxes, info = inputs_of(case)
        ch = make_chunker(case)
        res = ch.get_chunks(words, boxes, info, question_id=["q%d" % b for b in range(len(words))])
        res = json.loads(json.dumps(res))
        assert [crc(x) for x in res] == rec["crc"], case
        if "outputs" in rec:
            assert res == rec["outputs"]
        got = {k: {str(a): int(b) for a, b in v.items()} for k, v in ch.stats.items()}
        assert got == rec["stats"], case


@pytest.mark.parametrize("seed", [31, 32, 33])
def test_get_chunks_vs_oracle_larger(seed):
    from rag_docvqa_b200 import synth
    case = dict(seed=seed, docs=16, max_pages=12, max_words=700, max_layouts=30, clusters=seed % 2 == 1,
                cluster_layouts=seed % 2 == 1, chunk_size=60, overlap=10, tol=0.2, page_retrieval="concat")
    words, boxes, info = synth.make_chunker_batch(seed, 16, 12, 700, 30, clusters=case["clusters"])
    ch = make_chunker(case)
    got = ch.get_chunks(words, boxes, info, question_id=list(range(16)))
    want, stats = R.get_chunks(words, boxes, info, cluster_layouts=case["cluster_layouts"])
    assert json.loads(json.dumps(got)) == json.loads(json.dumps(want))
    assert dict(ch.stats["chunk_size_dist"]) == stats.chunk_size_dist
    assert dict(ch.stats["n_chunks_per_layout_dist"]) == stats.n_chunks_per_layout_dist


def test_layout_assign_bit_exact_on_arbitrary_floats():
    """Uniform random float64 boxes (products and quotients round): every decision and label equals Python's."""
    case = dict(chunk_size=60, overlap=10, tol=0.2, page_retrieval="concat", cluster_layouts=False)
    ch = make_chunker(case)
    rng = np.random.RandomState(4)
    pages, lays, labs = [], [], []
    for p in range(7):
        n, l = [0, 1, 31, 32, 33, 500, 1000][p], [3, 0, 5, 64, 1, 17, 9][p]
        x0, y0 = rng.rand(n), rng.rand(n)
        wb = np.stack([x0, y0, x0 + rng.rand(n) * 0.1, y0 + rng.rand(n) * 0.05], axis=1)
        # layout boxes cut through the words: many ratios near one half
        lx0, ly0 = rng.rand(l) * 0.5, rng.rand(l) * 0.5
        lb = np.stack([lx0, ly0, lx0 + rng.rand(l) * 0.5, ly0 + rng.rand(l) * 0.5], axis=1)
        for g in range(min(n, l)):                                  # box g covers half of word g, up to rounding
            lb[g] = [wb[g][0], wb[g][1] - 0.01, (wb[g][0] + wb[g][2]) / 2 + (g % 3 - 1) * 1e-17, wb[g][3] + 0.01]
        pages.append(wb); lays.append(lb); labs.append(np.arange(100, 100 + l, dtype=np.int32))
    inside, labels = ch.assign_words_to_layouts(pages, lays, labs, default_label=-7)
    near = 0
    for p in range(7):
        want = np.zeros((len(lays[p]), len(pages[p])), dtype=bool)
        want_label = np.full(len(pages[p]), -7, dtype=np.int32)
        for g, lbox in enumerate(lays[p].tolist()):
            for i, wbox in enumerate(pages[p].tolist()):
                r = R.containment_ratio(wbox, lbox)
                near += abs(r - 0.5) < 1e-3
                if r > 0.5:
                    want[g, i] = True; want_label[i] = 100 + g
        assert inside[p].shape == want.shape and np.array_equal(inside[p], want)
        assert np.array_equal(labels[p], want_label)
    assert near > 20


def test_integer_pixel_boxes_and_object_labels():
    """Boxes as ints (pixel coordinates, exact in float64) and labels that are not ints."""
    case = dict(chunk_size=10, overlap=2, tol=0.2, page_retrieval="concat", cluster_layouts=False)
    ch = make_chunker(case)
    rng = np.random.RandomState(9)
    n = 120
    x0, y0 = rng.randint(0, 800, size=n), rng.randint(0, 1000, size=n)
    boxes = [[[[int(a), int(b), int(a + w), int(b + 12)] for a, b, w in zip(x0, y0, rng.randint(1, 60, size=n))]]]
    words = [[["w%d" % i for i in range(n)]]]
    info = [[{"boxes": [[0, 0, 400, 500], [300, 300, 850, 1100], [0, 0, 400, 500]], "labels": ["title", "text", "table"]}]]
    got = ch.get_chunks(words, boxes, info, question_id=["q"])
    want, _ = R.get_chunks(words, boxes, info, chunk_size=10, overlap=2, tol=0.2)
    assert got == want
