"""Properties the oracle must have whatever the input (hypothesis; CPU).  The golden files pin the oracle to the reference on
seeded inputs; these pin the conventions the GPU tests rely on -- the ordering rules of the top-k, sharded == unsharded for
the merge, and the algebra by which csrc/vt5_embed.cu folds LayerNorm + Linear into tables -- on inputs nobody chose."""
import numpy as np
import torch
from hypothesis import given, settings, strategies as st

from oracle import ref_restated as R

SPECIALS = [0.0, -0.0, float("inf"), float("-inf"), float("nan"), 1.0, -1.0, 1e-38, -1e-38]
floats32 = st.one_of(st.sampled_from(SPECIALS), st.floats(width=32, allow_nan=True, allow_infinity=True),
                     st.integers(-3, 3).map(float))          # many exact duplicates


def rank_key(v: float, i: int):
    """(score desc, index asc) with torch.topk's conventions: NaN greatest, -0 == +0."""
    if np.isnan(v):
        return (0, 0.0, i)
    return (1, -(v + 0.0), i)                                # v + 0.0 turns -0.0 into +0.0


@settings(max_examples=200, deadline=None)
@given(st.lists(floats32, min_size=0, max_size=60), st.integers(1, 70))
def test_topk_lowest_index_is_the_stated_order(values, k):
    v = np.asarray(values, dtype=np.float32)
    got = R.topk_lowest_index(v, k).tolist()
    want = sorted(range(len(v)), key=lambda i: rank_key(float(v[i]), i))[:min(k, len(v))]
    assert got == want
    if len(v) and not np.isnan(v).any() and len(set(np.abs(v).tolist())) == len(v):      # distinct finite-or-inf scores
        assert got == R.topk_reference(torch.from_numpy(v), k).tolist()                   # == torch.topk, as the reference calls it


@settings(max_examples=200, deadline=None)
@given(st.lists(st.floats(width=32, allow_nan=False, allow_infinity=True), min_size=2, max_size=40))
def test_order_key_is_monotone(values):
    v = np.asarray(values, dtype=np.float32)
    key = R.order_key(v)
    for a in range(len(v)):
        for b in range(len(v)):
            if v[a] < v[b]:
                assert key[a] < key[b]
            elif v[a] == v[b]:
                assert key[a] == key[b]                       # includes -0.0 == +0.0
    assert R.order_key(np.asarray([np.nan], np.float32))[0] == 0xFFFFFFFF > key.max()


@settings(max_examples=100, deadline=None)
@given(st.integers(1, 5), st.integers(1, 4), st.integers(1, 12), st.integers(1, 12), st.randoms(use_true_random=False))
def test_merge_of_shard_topk_equals_topk_of_the_whole(world, n_q, per_shard, k, rnd):
    """The corpus-mode contract: every shard's local top-k (global ids), merged, is the top-k of the unsharded scores --
    for any shard sizes, with duplicates across shards, k larger than a shard, empty slots padded with id -1."""
    n = world * per_shard
    scores = np.asarray([[rnd.choice([-1.0, 0.0, 0.5, 1.0, rnd.random()]) for _ in range(n)] for _ in range(n_q)], np.float32)
    cand_s = np.full((n_q, world * k), -np.inf, np.float32)
    cand_i = np.full((n_q, world * k), -1, np.int64)
    for q in range(n_q):
        for r in range(world):
            lo = r * per_shard
            local = R.topk_lowest_index(scores[q, lo:lo + per_shard], k)
            cand_s[q, r * k:r * k + len(local)] = scores[q, lo + local]
            cand_i[q, r * k:r * k + len(local)] = lo + local
    got_s, got_i = R.merge_topk(cand_s, cand_i, k)
    for q in range(n_q):
        want = R.topk_lowest_index(scores[q], k)
        assert got_i[q, :len(want)].tolist() == want.tolist()
        assert (got_i[q, len(want):] == -1).all()
        np.testing.assert_array_equal(got_s[q, :len(want)], scores[q, want])


@settings(max_examples=100, deadline=None)
@given(st.lists(st.floats(-0.5, 1.5, width=32), min_size=4, max_size=4), st.integers(1, 3000), st.integers(1, 3000))
def test_crop_rectangle_is_ordered_and_truncates(bbox, w, h):
    x0, y0, x1, y1 = R.crop_rectangle(bbox, w, h)
    assert x0 <= x1 and y0 <= y1
    xs = sorted((int(bbox[0] * w), int(bbox[2] * w)))
    ys = sorted((int(bbox[1] * h), int(bbox[3] * h)))
    assert [x0, x1] == xs and [y0, y1] == ys                  # int() truncation, then the order fix (src/_modules.py:2108-2119)


@settings(max_examples=25, deadline=None)
@given(st.integers(1, 6).map(lambda n: 4 * n), st.integers(2, 40), st.integers(0, 2 ** 31 - 1))
def test_spatial_embedding_tables_are_the_reference_formula(D, n_pos, seed):
    """What csrc/vt5_embed.cu relies on: with x~ = x - mean(x) per row, XW = (x~ * gamma) W^T, YW likewise, c = beta W^T + b
    and the Gram tables of the centred rows, Linear(LayerNorm(x_l + y_u + x_r + y_b)) ==
    (XW[l] + YW[u] + XW[r] + YW[b]) / sqrt(var + eps) + c with var * D = ten Gram entries.  Float64 on both sides."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n_pos, D, generator=g, dtype=torch.float64) + 0.3
    y = torch.randn(n_pos, D, generator=g, dtype=torch.float64) * 0.7
    gamma = 1 + 0.2 * torch.randn(D, generator=g, dtype=torch.float64)
    beta = 0.1 * torch.randn(D, generator=g, dtype=torch.float64)
    W = torch.randn(D, D, generator=g, dtype=torch.float64) / D ** 0.5
    b = 0.1 * torch.randn(D, generator=g, dtype=torch.float64)
    eps = 1e-12
    bbox = torch.randint(0, n_pos, (2, 9, 4), generator=g)
    bbox[0, 0] = 0
    bbox[0, 1] = torch.tensor([3 % n_pos, 3 % n_pos, 3 % n_pos, 3 % n_pos])
    ref = R.spatial_embeddings(bbox, x, y, gamma, beta, eps, W, b)
    xc, yc = x - x.mean(1, keepdim=True), y - y.mean(1, keepdim=True)
    XW, YW = (xc * gamma) @ W.T, (yc * gamma) @ W.T
    gxx, gxy, gyy = xc @ xc.T, xc @ yc.T, yc @ yc.T
    c = beta @ W.T + b
    l, u, r, bb = (bbox[..., i] for i in range(4))
    var = (gxx[l, l] + gyy[u, u] + gxx[r, r] + gyy[bb, bb]
           + 2 * (gxy[l, u] + gxx[l, r] + gxy[l, bb] + gxy[r, u] + gyy[u, bb] + gxy[r, bb])) / D
    got = (XW[l] + YW[u] + XW[r] + YW[bb]) / torch.sqrt(var.clamp(min=0) + eps)[..., None] + c
    # where the four rows nearly cancel (var ~ eps) LayerNorm itself is ill-conditioned: compare where it is not
    ok = var > 1e-6
    torch.testing.assert_close(got[ok], ref[ok], rtol=1e-8, atol=1e-8)
