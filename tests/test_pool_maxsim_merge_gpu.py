"""GPU parity: mean pooling, MaxSim late interaction, stand-alone top-k and shard merge vs the oracle."""
import os

import numpy as np
import pytest
import torch

from oracle import compare
from oracle import ref_restated as R
from rag_docvqa_b200 import synth

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

# fp32 tolerances, stated: pooling and MaxSim differ from torch only in summation order.
POOL_RTOL, POOL_ATOL = 1e-5, 1e-6
MAXSIM_RTOL = 1e-5          # north_star: scores within 1e-5 relative


def test_mean_pooling_golden(golden_dir):
    from rag_docvqa_b200 import functional as F
    z = np.load(os.path.join(golden_dir, "mean_pooling.npz"))
    got = F.mean_pooling(torch.from_numpy(z["embs"]).to(DEV), torch.from_numpy(z["mask"]).to(DEV))
    np.testing.assert_allclose(got.cpu().numpy(), z["pooled"], rtol=POOL_RTOL, atol=POOL_ATOL)
    assert (got[:2] == 0).all()          # all-pad rows -> exactly 0 (clamp 1e-9)


@pytest.mark.parametrize("n,d,mean_len", [(64, 384, 96), (33, 768, 40), (20, 1024, 200), (9, 100, 12),
                                          (5, 2048, 30), (40, 64, 8)])
def test_mean_pooling_shapes(n, d, mean_len):
    from rag_docvqa_b200 import functional as F
    embs, mask = synth.make_token_batch(n, d, 11 + n, mean_len=mean_len, std_len=mean_len / 4, min_len=0,
                                        max_len=512, all_pad_rows=1)
    ref = R.mean_pooling(embs, mask)
    got = F.mean_pooling(embs.to(DEV), mask.to(DEV))
    np.testing.assert_allclose(got.cpu().numpy(), ref.numpy(), rtol=POOL_RTOL, atol=POOL_ATOL)


def test_mean_pooling_weighted_mask_and_normalise():
    from rag_docvqa_b200 import functional as F
    g = torch.Generator().manual_seed(2)
    embs = torch.randn(12, 50, 384, generator=g)
    mask = torch.randint(0, 3, (12, 50), generator=g)          # the reference multiplies by the mask VALUE
    ref = R.mean_pooling(embs, mask)
    got, bf, nrm = F.mean_pooling(embs.to(DEV), mask.to(DEV), out_bf16=True, return_norm=True)
    np.testing.assert_allclose(got.cpu().numpy(), ref.numpy(), rtol=POOL_RTOL, atol=POOL_ATOL)
    np.testing.assert_allclose(nrm.cpu().numpy(), ref.norm(dim=-1).numpy(), rtol=1e-5)
    assert torch.equal(bf.cpu(), got.cpu().to(torch.bfloat16))
    gotn = F.mean_pooling(embs.to(DEV), mask.to(DEV), normalise=True)
    refn = torch.nn.functional.normalize(ref, p=2, dim=-1)
    np.testing.assert_allclose(gotn.cpu().numpy(), refn.numpy(), rtol=1e-5, atol=1e-6)


MAXSIM_MODES = ["ffma", "tf32x3", "auto"]      # CUDA-core fp32, 3xTF32 on the tensor pipe, the shipped default


@pytest.mark.parametrize("mode", MAXSIM_MODES)
def test_late_interaction_golden(golden_dir, mode):
    from rag_docvqa_b200 import functional as F
    z = np.load(os.path.join(golden_dir, "late_interaction.npz"))
    q = torch.from_numpy(z["q"]).to(DEV)
    for b, (pk, sk) in enumerate((("p0", "s0"), ("p1", "s1"))):
        got = F.late_interaction(q[b:b + 1], torch.from_numpy(z[pk]).to(DEV), mode=mode)
        np.testing.assert_allclose(got.cpu().numpy(), z[sk], rtol=MAXSIM_RTOL)


@pytest.mark.parametrize("mode", MAXSIM_MODES)
@pytest.mark.parametrize("n,Lq,Lp,d", [(3, 128, 128, 64), (5, 200, 77, 96), (2, 2048, 2048, 768), (7, 1, 300, 768),
                                       (4, 130, 1, 36), (1, 257, 513, 20), (50, 300, 700, 768), (3, 64, 64, 2048)])
def test_late_interaction_shapes(n, Lq, Lp, d, mode):
    from rag_docvqa_b200 import functional as F
    g = torch.Generator().manual_seed(n * 1000 + Lq)
    q = torch.randn(1, Lq, d, generator=g)
    p = torch.randn(n, Lp, d, generator=g)
    p[0, 0] = 0.0                                               # zero token: F.normalize eps path
    ref64 = R.late_interaction_f64(q, p).numpy()
    ref32 = R.late_interaction(q, p).numpy()
    got = F.late_interaction(q.to(DEV), p.to(DEV), mode=mode).cpu().numpy()
    np.testing.assert_allclose(got, ref64, rtol=MAXSIM_RTOL)
    if mode == "ffma":
        # no further from the float64 truth than torch's own fp32 result is (x4 slack)
        assert np.abs(got - ref64).max() <= 4 * max(np.abs(ref32 - ref64).max(), 1e-6 * np.abs(ref64).max())
    else:
        # 3xTF32: exact products, but the tensor core's accumulator rounds toward zero -> ~d * 2.1e-9 relative low
        # (stated for sums of positive maxima; with Lp == 1 the terms change sign and cancel, only the 1e-5 bar applies)
        if Lp >= 64:
            assert np.abs(got / ref64 - 1).max() <= max(1e-6, d * 3.2e-9)


def test_split_tf32_is_exact():
    """hi + lo == x bit for bit, hi has a 10-bit mantissa, |lo| <= 2^-11 |x| (the premise of the 3xTF32 mode)."""
    from rag_docvqa_b200 import functional as F
    g = torch.Generator().manual_seed(4)
    x = torch.randn(37, 100, generator=g) * torch.logspace(-20, 20, 37).unsqueeze(1)
    hi, lo = F.split_tf32(x.to(DEV))
    hi, lo = hi.cpu(), lo.cpu()
    assert torch.equal(hi + lo, x)
    assert (hi.view(torch.int32) & 0x1FFF).eq(0).all()
    assert (lo.abs() <= x.abs() * 2.0 ** -11).all()
    hn, ln = F.split_tf32(x.to(DEV), normalise=True)
    ref = torch.nn.functional.normalize(x, p=2, dim=-1)
    np.testing.assert_allclose((hn.cpu().double() + ln.cpu().double()).numpy(), ref.double().numpy(), rtol=1e-6, atol=1e-30)


def test_topk_segments_matches_oracle():
    from rag_docvqa_b200 import functional as F
    g = torch.Generator().manual_seed(8)
    scores = [torch.randn(n, generator=g) for n in (50, 0, 3, 20000, 1)]
    scores[0][7] = scores[0][3]
    scores[3][100:200] = 5.0
    idx, val, cnt = F.topk_segments([s.to(DEV) for s in scores], 10)
    idx, cnt = idx.cpu().numpy(), cnt.cpu().numpy()
    for b, s in enumerate(scores):
        kb = min(10, len(s))
        assert cnt[b] == kb
        np.testing.assert_array_equal(idx[b, :kb], R.topk_lowest_index(s, 10))
        assert (idx[b, kb:] == -1).all()


@pytest.mark.parametrize("world,k", [(2, 10), (8, 10), (4, 5)])
def test_topk_merge_equals_unsharded(world, k):
    from rag_docvqa_b200 import functional as F
    g = torch.Generator().manual_seed(world)
    Q, n = 37, 4000
    scores = torch.randn(Q, n, generator=g)
    scores[:, 100] = scores[:, 3000]                            # cross-shard exact ties
    shard = n // world
    cv, ci = [], []
    for r in range(world):
        part = scores[:, r * shard:(r + 1) * shard]
        loc = np.stack([R.topk_lowest_index(part[qi], k) for qi in range(Q)])
        ci.append(torch.from_numpy(loc + r * shard))
        cv.append(torch.gather(part, 1, torch.from_numpy(loc)))
    cand_v, cand_i = torch.cat(cv, 1), torch.cat(ci, 1)
    cand_i[0, 3] = -1                                           # an empty slot is skipped
    out_v, out_i = F.topk_merge(cand_v.to(DEV), cand_i.to(DEV), k)
    ref_v, ref_i = R.merge_topk(cand_v.numpy(), cand_i.numpy(), k)
    np.testing.assert_array_equal(out_i.cpu().numpy(), ref_i)
    np.testing.assert_array_equal(out_v.cpu().numpy(), ref_v)
    for qi in range(1, Q):                                      # equals the unsharded selection
        np.testing.assert_array_equal(out_i[qi].cpu().numpy(), R.topk_lowest_index(scores[qi], k))


def test_pooled_patch_golden(golden_dir):
    """pooled_patch_topk (pool -> cosine of every patch vector -> top-k patches, strip max -> top-k strips) against what the
    reference's own mean_pooling + Retriever._get_similarities + torch.max / torch.topk produced."""
    from rag_docvqa_b200 import functional as F
    z = np.load(os.path.join(golden_dir, "pooled_patch.npz"))
    n_docs, k = int(z["n_docs"]), int(z["k"])
    patches = [torch.from_numpy(z["patches_%d" % b]).to(DEV) for b in range(n_docs)]
    q, mask = torch.from_numpy(z["q"]).to(DEV), torch.from_numpy(z["mask"]).to(DEV)
    res = F.pooled_patch_topk(patches, q, k, question_mask=mask)
    torch.cuda.synchronize()
    np.testing.assert_allclose(res.question.cpu().numpy(), z["pooled"], rtol=1e-5, atol=1e-6)
    p_idx, p_cnt = res.patch_idx.cpu().numpy(), res.patch_cnt.cpu().numpy()
    s_idx, s_cnt = res.strip_idx.cpu().numpy(), res.strip_cnt.cpu().numpy()
    for b in range(n_docs):
        ref, own = z["sims_%d" % b], res.similarities[b].cpu().numpy()
        compare.assert_scores_close(own, ref, what="doc %d patch sims" % b)
        kb = min(k, len(ref))
        assert p_cnt[b] == kb and (p_idx[b, kb:] == -1).all()
        compare.assert_topk_matches(p_idx[b, :kb], own, ref, k, what="doc %d patches" % b)
        sref, sown = z["strip_%d" % b], res.strip_scores[b].cpu().numpy()
        compare.assert_scores_close(sown, sref, what="doc %d strip scores" % b)
        n = len(sref)
        if n:
            np.testing.assert_array_equal(sown, own.reshape(n, -1).max(axis=1))          # exactly the best patch of the strip
        ks = min(k, n)
        assert s_cnt[b] == ks and (s_idx[b, ks:] == -1).all()
        compare.assert_topk_matches(s_idx[b, :ks], sown, sref, k, what="doc %d strips" % b)


def test_pooled_patch_c4_shape_and_nan():
    """One document at C4 strip shape (8 strips x 2048 patches x 768) against the oracle; a NaN patch makes its strip NaN
    (torch.max) and therefore the best strip (torch.topk: NaN greatest)."""
    from rag_docvqa_b200 import functional as F
    g = torch.Generator().manual_seed(4)
    patches = [torch.randn(8, 2048, 768, generator=g), torch.randn(3, 2048, 768, generator=g)]
    patches[1][2, 100, 5] = float("nan")
    q = torch.randn(2, 2048, 768, generator=g)
    sims, strips, _ = R.pooled_patch_scores(patches, q)
    res = F.pooled_patch_topk([p.to(DEV) for p in patches], q.to(DEV), 5, k_strips=2)
    for b in range(2):
        own = res.similarities[b].cpu().numpy()
        compare.assert_scores_close(own, sims[b].numpy())
        compare.assert_topk_matches(res.patch_idx[b].cpu().numpy()[:5], own, sims[b].numpy(), 5)
    assert np.isnan(res.strip_scores[1].cpu().numpy()[2]) and res.strip_idx.cpu().numpy()[1, 0] == 2
    compare.assert_scores_close(res.strip_scores[0].cpu().numpy(), strips[0].numpy())
    assert res.strip_idx.shape == (2, 2)


def test_visual_retriever_pooled_score():
    """VisualRetriever(config, score='pooled'): the strips the decode receives are ranked by the pooled-patch score."""
    from PIL import Image
    from rag_docvqa_b200.retriever import VisualRetriever
    g = torch.Generator().manual_seed(9)
    n, L, d = 6, 64, 32
    patches = [torch.randn(n, L, d, generator=g)]
    q = torch.randn(1, 16, d, generator=g)
    _, strips, _ = R.pooled_patch_scores(patches, q)
    want = sorted(R.topk_lowest_index(strips[0], 2).tolist())
    pages = [[Image.new("RGB", (40, 30), (i, 0, 0)) for i in range(n)]]
    vr = VisualRetriever({"chunk_num": 2, "include_surroundings": 0, "chunk_mode": "horizontal", "device": DEV}, score="pooled")
    crops, page_ids = vr.retrieve([p.to(DEV) for p in patches], q.to(DEV), [np.arange(n)], [[[[pages[0][i]]] for i in range(n)]],
                                  [[[[0, 0, 40, 30]] for i in range(n)]], pages)
    assert page_ids[0] == want and len(crops[0]) == 2


def test_late_interaction_auto_at_the_widest_supported_embedding():
    """mode='auto' picks the 3xTF32 tensor-core kernel up to d = 2048, where its systematic error is largest (about
    d * 2.1e-9 relative low: 4.4e-6): held to north_star's 1e-5 against float64, at Lq = Lp = 2048."""
    from rag_docvqa_b200 import functional as F
    g = torch.Generator().manual_seed(8)
    q = torch.randn(1, 2048, 2048, generator=g)
    p = torch.randn(2, 2048, 2048, generator=g)
    got = F.late_interaction(q.to(DEV), p.to(DEV), mode="auto").cpu().numpy().astype(np.float64)
    ref = R.late_interaction_f64(q, p).numpy()
    rel = np.abs(got - ref) / np.abs(ref)
    assert rel.max() < 1e-5, rel
    strict = F.late_interaction(q.to(DEV), p.to(DEV), mode="ffma").cpu().numpy().astype(np.float64)
    assert (np.abs(strict - ref) / np.abs(ref)).max() < 2e-6
    assert (got <= strict + 1e-3).all()                       # the tensor-core accumulator rounds toward zero: never high


def test_mean_pooling_masked_nan_is_not_read():
    """Documented deviation: a NaN / Inf at a MASKED token position is ignored -- the kernel never
    reads masked tokens -- where the reference's `embs * mask` turns it into NaN for the whole chunk.  Pinned both ways."""
    from rag_docvqa_b200 import functional as F
    embs, mask = synth.make_token_batch(12, 64, 3, mean_len=12.0, std_len=4.0, min_len=3, max_len=20)
    dirty = embs.clone()
    row = int((mask.sum(dim=1) < mask.shape[1]).nonzero()[0])           # a chunk that has padding
    first_pad = int(mask[row].sum())
    dirty[row, first_pad, 5] = float("nan")
    dirty[row, -1, 7] = float("inf")
    ref_clean = R.mean_pooling(embs, mask)
    ref_dirty = R.mean_pooling(dirty, mask)
    assert torch.isnan(ref_dirty[row]).any()                             # the reference propagates it ...
    got = F.mean_pooling(dirty.to(DEV), mask.to(DEV)).cpu()
    assert torch.isfinite(got).all()                                     # ... the kernel does not read it
    np.testing.assert_allclose(got.numpy(), ref_clean.numpy(), rtol=1e-5, atol=1e-6)
    unmasked = dirty.clone()
    unmasked[row, 0, 5] = float("nan")                                   # an UNMASKED NaN is data: it propagates here too
    assert torch.isnan(F.mean_pooling(unmasked.to(DEV), mask.to(DEV)).cpu()[row, 5])


@pytest.mark.parametrize("split", ["auto", "0", "2", "16"])
@pytest.mark.parametrize("n,L,d", [(8, 2048, 768), (3, 1000, 1024), (1, 4096, 64), (5, 70, 768)])
def test_mean_pooling_row_split_over_a_cluster(monkeypatch, split, n, L, d):
    """Few rows of many tokens (the rendered questions of pooled-patch retrieval): a row is summed by a thread-block cluster.
    Every split factor gives the oracle's result within fp32 summation error, bit-identically from run to run, and the
    normalised / bf16 / norm outputs agree with it."""
    from rag_docvqa_b200 import functional as F
    if split == "auto":
        monkeypatch.delenv("RDV_POOL_SPLIT", raising=False)
    else:
        monkeypatch.setenv("RDV_POOL_SPLIT", split)
    g = torch.Generator().manual_seed(n * 1000 + L)
    embs = torch.randn(n, L, d, generator=g)
    mask = (torch.rand(n, L, generator=g) < 0.8).to(torch.int64)
    mask[0, L // 2:] = 0                                        # a row whose second half is padding: whole CTAs see no token
    if n > 1:
        mask[1] = 0                                             # nothing at all: clamp(min=1e-9) keeps 0 / 1e-9 = 0
    ref = R.mean_pooling(embs, mask)
    e, m = embs.to(DEV), mask.to(DEV)
    got, bf, nrm = F.mean_pooling(e, m, out_bf16=True, return_norm=True)
    again = F.mean_pooling(e, m)
    assert torch.equal(got, again)
    torch.testing.assert_close(got.cpu(), ref, rtol=2e-5, atol=2e-6)
    torch.testing.assert_close(nrm.cpu(), ref.norm(dim=1), rtol=2e-5, atol=2e-6)
    assert torch.equal(bf.cpu(), got.cpu().to(torch.bfloat16))
    gotn = F.mean_pooling(e, m, normalise=True)
    torch.testing.assert_close(gotn.cpu(), torch.nn.functional.normalize(ref, dim=1), rtol=2e-5, atol=2e-6)


def test_pooled_patch_fuzz():
    """Twelve seeded batches nobody chose: 1-6 documents of 0-9 strips, L in {1, 3, 17, 64, 300}, k and k_strips from 1 to
    12 (above and below L and the strip count), a padded question mask, duplicated patches (exact ties) -- against the
    oracle's composition of the reference's functions."""
    from rag_docvqa_b200 import functional as F
    rng = np.random.RandomState(31)
    for case in range(12):
        B = int(rng.randint(1, 7))
        L = int(rng.choice([1, 3, 17, 64, 300]))
        d = 4 * int(rng.choice([2, 24, 96, 192]))
        k, ks = int(rng.randint(1, 13)), int(rng.randint(1, 13))
        Lq = int(rng.randint(1, 40))
        g = torch.Generator().manual_seed(500 + case)
        u = torch.randn(d, generator=g)
        patches = [torch.randn(int(rng.randint(0, 10)), L, d, generator=g) + 0.5 * u for _ in range(B)]
        for p in patches:
            if p.shape[0] >= 2 and L >= 2:
                p[1, L - 1] = p[0, 0]                             # an exact tie across strips: lowest index first
        q = torch.randn(B, Lq, d, generator=g) + 0.5 * u
        mask = (torch.rand(B, Lq, generator=g) < 0.7).to(torch.int64)
        mask[:, 0] = 1
        sims, strips, pooled = R.pooled_patch_scores(patches, q, mask)
        res = F.pooled_patch_topk([p.to(DEV) for p in patches], q.to(DEV), k, question_mask=mask.to(DEV), k_strips=ks)
        torch.cuda.synchronize()
        p_idx, p_cnt = res.patch_idx.cpu().numpy(), res.patch_cnt.cpu().numpy()
        s_idx, s_cnt = res.strip_idx.cpu().numpy(), res.strip_cnt.cpu().numpy()
        for b in range(B):
            what = "case %d doc %d (L=%d k=%d ks=%d)" % (case, b, L, k, ks)
            ref, own = sims[b].numpy(), res.similarities[b].cpu().numpy()
            compare.assert_scores_close(own, ref, what=what)
            kb = min(k, len(ref))
            assert p_cnt[b] == kb and (p_idx[b, kb:] == -1).all(), what
            compare.assert_topk_matches(p_idx[b, :kb], own, ref, k, what=what)
            n = patches[b].shape[0]
            sown = res.strip_scores[b].cpu().numpy()
            if n:
                np.testing.assert_array_equal(sown, own.reshape(n, -1).max(axis=1))
            kb = min(ks, n)
            assert s_cnt[b] == kb and (s_idx[b, kb:] == -1).all(), what
            compare.assert_topk_matches(s_idx[b, :kb], sown, strips[b].numpy(), ks, what=what)
