"""TEST INFRASTRUCTURE ONLY -- comparators implementing SURVEY.md section 8c's parity rules.

fp32 mode (north_star): top-k indices bit-exact except across ties within 1e-6; scores within
1e-5 relative (with the same 1e-6 absolute floor the tie rule uses, because a cosine that is the
difference of two near-cancelling sums has no meaningful relative error).
"""
from __future__ import annotations

import numpy as np

SCORE_RTOL = 1e-5
SCORE_ATOL = 1e-6
TIE_TOL = 1e-6


def assert_scores_close(got, ref, rtol: float = SCORE_RTOL, atol: float = SCORE_ATOL, what: str = "scores"):
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    assert got.shape == ref.shape, "%s: shape %s != %s" % (what, got.shape, ref.shape)
    if got.size == 0:
        return
    both_nan = np.isnan(got) & np.isnan(ref)
    err = np.abs(got - ref)
    bound = atol + rtol * np.abs(ref)
    bad = ~both_nan & ~(err <= bound)
    assert not bad.any(), "%s: %d/%d outside |d|<=%.0e+%.0e|ref|; worst %.3e at %s (got %r ref %r)" % (
        what, int(bad.sum()), got.size, atol, rtol, float(np.nanmax(np.where(bad, err, 0))),
        np.argwhere(bad)[0].tolist(), got[tuple(np.argwhere(bad)[0])], ref[tuple(np.argwhere(bad)[0])])


def assert_topk_matches(got_idx, got_scores, ref_scores, k: int, tie_tol: float = TIE_TOL, what: str = "topk"):
    """got_idx: the new path's hits in rank order for ONE document; got_scores: the new path's own
    full score vector; ref_scores: the oracle's full score vector for the same document.

    1. count: min(k, n) hits, all distinct, all in range;
    2. against the oracle: every returned index has an oracle score >= (oracle k-th best - tie_tol),
       and every index the oracle ranks strictly above (k-th best + tie_tol) is returned
       ("identical index sets modulo tie groups");
    3. order: oracle scores along the returned ranking are non-increasing within tie_tol;
    4. determinism on the path's OWN scores (bit-exact rule): the ranking is descending by
       (own score, then lowest index), and no excluded index beats or ties-with-lower-index the last hit.
    """
    got_idx = np.asarray(got_idx, dtype=np.int64)
    ref = np.asarray(ref_scores, dtype=np.float64)
    own = np.asarray(got_scores, dtype=np.float32)
    n = ref.shape[0]
    k_min = min(k, n)
    assert got_idx.shape[0] == k_min, "%s: %d hits, expected %d" % (what, got_idx.shape[0], k_min)
    if k_min == 0:
        return
    assert len(set(got_idx.tolist())) == k_min, "%s: duplicate hits %s" % (what, got_idx)
    assert got_idx.min() >= 0 and got_idx.max() < n, "%s: index out of range %s" % (what, got_idx)
    ref_rank = np.where(np.isnan(ref), np.inf, ref)      # NaN sorts greatest (torch.topk)
    kth = np.sort(ref_rank)[::-1][k_min - 1]
    assert (ref_rank[got_idx] >= kth - tie_tol).all(), "%s: hit below the oracle's k-th score: %s" % (
        what, got_idx[ref_rank[got_idx] < kth - tie_tol])
    must = np.nonzero(ref_rank > kth + tie_tol)[0]
    missing = set(must.tolist()) - set(got_idx.tolist())
    assert not missing, "%s: oracle-certain hits missing: %s" % (what, sorted(missing))
    seq = ref_rank[got_idx]
    finite = np.isfinite(seq[:-1]) | np.isfinite(seq[1:])
    assert ((seq[:-1] >= seq[1:] - tie_tol) | ~finite).all(), "%s: not descending vs oracle: %s" % (what, seq)
    # determinism on own scores
    from oracle.ref_restated import topk_lowest_index
    expect = topk_lowest_index(own, k)
    assert np.array_equal(expect, got_idx), "%s: not (score desc, index asc) on own scores: got %s expect %s" % (
        what, got_idx, expect)


def recall_at_k(got_idx, ref_idx) -> float:
    got_idx = np.asarray(got_idx)
    ref_idx = np.asarray(ref_idx)
    hits, total = 0, 0
    for g, r in zip(got_idx, ref_idx):
        r = set(int(x) for x in r if x >= 0)
        hits += len(r & set(int(x) for x in g if x >= 0))
        total += len(r)
    return hits / max(1, total)


def assert_order_matches_modulo_ties(got, ref, scores):
    """Reranker index lists: same length, distinct entries, and the same score at every rank (equal scores may
    resolve differently: the reference's order of ties depends on which numpy sort kernel the CPU dispatches to)."""
    assert len(got) == len(ref) and len(set(got)) == len(got)
    assert [float(scores[i]) for i in got] == [float(scores[i]) for i in ref], (got, ref)


def load_postproc_golden(golden_dir):
    """tests/golden/postproc.json (oracle/make_golden_postproc.py) with the packed arrays decoded."""
    import base64
    import json
    import os
    with open(os.path.join(golden_dir, "postproc.json")) as f:
        g = json.load(f)
    for c in g["page_vote"]:
        c["sims"] = [np.frombuffer(base64.b64decode(x), dtype=np.float32) for x in c["sims_f32_b64"]]
    for c in g["rerank"]:
        c["scores_np"] = np.asarray(c["scores"], dtype=np.float32 if c["dtype"] == "f32" else np.float64)
    return g
