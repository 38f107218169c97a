"""Oracle restatement of the reranker post-processing and the page vote against the
reference's frozen outputs (tests/golden/postproc.json, oracle/make_golden_postproc.py) and live CPython."""
import random

import numpy as np

from oracle import ref_restated as R
from oracle.compare import assert_order_matches_modulo_ties, load_postproc_golden as load_postproc


def test_rerank_order_matches_reference(golden_dir):
    g = load_postproc(golden_dir)
    exact = 0
    for c in g["rerank"]:
        got = R.rerank_order(c["scores_np"], c["thresh"], c["max"], c["min"])
        assert_order_matches_modulo_ties(got, c["order"], c["scores_np"])
        exact += got == c["order"]
        if len(set(c["scores"])) == len(c["scores"]):
            assert got == c["order"]
    assert exact >= len(g["rerank"]) * 0.8


def test_rerank_conventions():
    s = np.array([0.5, np.nan, 0.5, 0.9, 0.1], dtype=np.float32)
    assert R.rerank_order(s, 0.4, 5, 1) == [3, 2, 0]              # NaN filtered; ties: higher index first
    assert R.rerank_order(s, 0.95, 5, 2) == [1, 3]                # fallback takes the head of the sort, NaN first
    assert R.rerank_order(s, 0.0, 2, 1) == [3, 2]
    assert R.rerank_order(np.zeros(0, dtype=np.float32), 0.4, 5, 1) == []
    cands, ids = R.rerank(s, list("abcde"), [10, 11, 12, 13, 14])
    assert cands == ["d", "c", "a"] and ids == [13, 12, 10]


def test_int_set_order_is_cpython_set_order():
    rnd = random.Random(5)
    for _ in range(20000):
        n = rnd.randint(0, 64)
        hi = rnd.choice([3, 8, 20, 50, 200, 1000, 100000])
        v = [rnd.randint(0, hi) for _ in range(n)]
        assert R.int_set_order(v) == list(set(v))


def test_page_vote_matches_reference(golden_dir):
    g = load_postproc(golden_dir)
    n = 0
    for c in g["page_vote"]:
        for b, pages in enumerate(c["pages"]):
            s = c["sims"][b]
            got = R.page_vote(pages, s, len(s), c["mode"] == "weightmajorpage", legacy_promotion=False)
            assert got == c["major"][b], (c["mode"], pages)
            n += 1
    assert n >= 200


def test_page_vote_conventions():
    assert R.page_vote([], np.zeros(0, np.float32), 0, False) == 0
    # the weights come from the first k CHUNKS, not from the hits (src/RAGVT5.py:468: zip(pages, weights))
    sims = np.array([0.1, 0.1, 0.9, 0.0, 0.0], dtype=np.float32)
    assert R.page_vote([7, 7, 3], sims, 5, True) == 3
    assert R.page_vote([7, 7, 3], sims, 5, False) == 7
    # tie between pages: first in CPython's set order (8 -> slot 0, 1 -> slot 1), not first seen
    assert R.page_vote([1, 8], sims, 5, False) == 8 and list(set([1, 8]))[0] == 8
    # a negative sum flips the vote
    assert R.page_vote([0, 1], np.array([0.5, 0.1, -2.0], np.float32), 3, True) == 1
