"""GPU parity: fused cosine score + segmented top-k (librdv rdv_score_topk_f32) vs the oracle.

fp32 mode bars (north_star): identical top-k index sets modulo ties within 1e-6 (checked by
oracle/compare.py, plus bit-exact (score desc, index asc) order on the kernel's own scores);
scores within 1e-5 relative (+1e-6 absolute floor)."""
import os

import numpy as np
import pytest
import torch

from oracle import compare
from oracle import ref_restated as R
from rag_docvqa_b200 import synth

pytestmark = pytest.mark.gpu

TEXT_CASES = ["c1", "ragged_norm", "ragged_raw", "k20_d1024"]


LDG, TMA, CLUSTER = 1, 2, 3         # CLUSTER = the one-launch kernel (a thread-block cluster per document)
ALGOS = [LDG, TMA, CLUSTER]


def run_case(emb, q, k, tile_rows=0, algo=0):
    """algo 0: what the product picks (rdv_score_plan + rdv_retrieve_plan); LDG / TMA: the two-launch path with that
    streaming kernel; CLUSTER: the one-launch cluster kernel where its limits allow (k <= 32, documents <= 2048 rows),
    else the product's choice."""
    from rag_docvqa_b200 import _lib
    from rag_docvqa_b200 import functional as F
    dev = torch.device("cuda:0")
    cluster = None if algo == 0 else False
    if algo == CLUSTER:
        fits = k <= _lib.lib.rdv_cluster_max_k() and max([e.shape[0] for e in emb] + [0]) <= _lib.lib.rdv_cluster_max_rows(8)
        cluster, algo = (True if fits else None), 0
    res = F.score_topk([e.to(dev) for e in emb], q.to(dev), k, tile_rows=tile_rows, algo=algo, cluster=cluster)
    torch.cuda.synchronize()
    return res


def check_against_oracle(res, emb, q, k, ref_sims=None):
    ref = ref_sims if ref_sims is not None else [s.numpy() for s in R.score(emb, q)]
    idx = res.topk_idx.cpu().numpy()
    val = res.topk_val.cpu().numpy()
    cnt = res.topk_cnt.cpu().numpy()
    assert len(res.similarities) == len(emb)
    for b in range(len(emb)):
        n = emb[b].shape[0]
        own = res.similarities[b].cpu().numpy()
        assert own.shape == (n,)
        compare.assert_scores_close(own, ref[b], what="doc %d sims" % b)
        kb = min(k, n)
        assert cnt[b] == kb
        assert (idx[b, kb:] == -1).all() and np.isneginf(val[b, kb:]).all()
        compare.assert_topk_matches(idx[b, :kb], own, ref[b], k, what="doc %d" % b)
        np.testing.assert_array_equal(val[b, :kb], own[idx[b, :kb]])


@pytest.mark.parametrize("algo", ALGOS)
@pytest.mark.parametrize("name", TEXT_CASES)
def test_golden_cases(golden_dir, name, algo):
    z = np.load(os.path.join(golden_dir, "score_topk_%s.npz" % name))
    sizes = z["sizes"].tolist()
    k = int(z["k"])
    emb = [torch.from_numpy(z["emb_%d" % b]) for b in range(len(sizes))]
    q = torch.from_numpy(z["q"])
    ref = [z["sims_%d" % b] for b in range(len(sizes))]
    res = run_case(emb, q, k, algo=algo)
    check_against_oracle(res, emb, q, k, ref_sims=ref)   # the reference's own frozen outputs


@pytest.mark.parametrize("algo,tile_rows", [(LDG, 8), (LDG, 32), (LDG, 128), (CLUSTER, 0), (TMA, 1), (TMA, 5), (TMA, 8), (TMA, 16)])
@pytest.mark.parametrize("normalised", [True, False])
def test_c2_full(algo, tile_rows, normalised):
    batch = synth.make_text_batch("C2", normalised=normalised)
    res = run_case(batch["text_embeddings"], batch["question_embeddings"], 5, tile_rows=tile_rows, algo=algo)
    check_against_oracle(res, batch["text_embeddings"], batch["question_embeddings"], 5)


@pytest.mark.parametrize("algo", ALGOS)
@pytest.mark.parametrize("k", [1, 5, 10, 20, 64])
def test_k_sweep_with_duplicates(k, algo):
    sizes = [300, 17, 0, 64, 1, 1000, 3000, 4097, 8000]   # register (<=1024, <=4096), cached and L2 selection paths
    if algo == CLUSTER:
        sizes = [300, 17, 0, 64, 1, 1000, 1024, 1023, 33, 31, 32, 256, 257, 512, 513, 511]   # slices of <= 32 / 64 / 128 rows, edges
    emb, q = synth.make_embeddings(sizes, 384, 31 + k, dup_frac=0.2)
    res = run_case(emb, q, k, algo=algo)
    check_against_oracle(res, emb, q, k)


@pytest.mark.parametrize("d", [4, 64, 100, 128, 256, 384, 512, 640, 768, 1024, 2048])
def test_dims(d):
    sizes = [50, 3, 129]
    emb, q = synth.make_embeddings(sizes, d, 7 + d, normalised=False, dup_frac=0.05)
    res = run_case(emb, q, 5)
    check_against_oracle(res, emb, q, 5)
    if d in (128, 256, 384, 512, 768, 1024):
        for algo in ALGOS:
            res = run_case(emb, q, 5, algo=algo)
            check_against_oracle(res, emb, q, 5)
    else:
        from rag_docvqa_b200 import _lib
        with pytest.raises(_lib.RdvError):
            run_case(emb, q, 5, algo=TMA)


def test_special_values():
    d = 128
    g = torch.Generator().manual_seed(3)
    e = torch.randn(40, d, generator=g)
    e[3] = 0.0                      # zero chunk scores exactly 0, never NaN  
    e[7] = e[5]                     # exact duplicate: lower index first
    e[11, 0] = float("nan")         # NaN sorts greatest (torch.topk convention)
    e[20] = -e[5]
    q = torch.randn(1, d, generator=g)
    res = run_case([e], q, 6)
    own = res.similarities[0].cpu().numpy()
    assert own[3] == 0.0
    assert np.isnan(own[11])
    idx = res.topk_idx.cpu().numpy()[0]
    assert idx[0] == 11
    ref = R.score([e], q)[0].numpy()
    compare.assert_scores_close(own, ref)
    np.testing.assert_array_equal(idx, R.topk_lowest_index(own, 6))
    pos5, pos7 = list(idx).index(5) if 5 in idx else None, list(idx).index(7) if 7 in idx else None
    if pos5 is not None and pos7 is not None:
        assert pos5 + 1 == pos7


def test_all_empty_and_zero_question():
    emb = [torch.zeros(0, 384), torch.zeros(0, 384)]
    q = torch.randn(2, 384)
    res = run_case(emb, q, 5)
    assert res.topk_cnt.cpu().tolist() == [0, 0]
    assert (res.topk_idx.cpu().numpy() == -1).all()
    emb = [torch.randn(10, 384)]
    res = run_case(emb, torch.zeros(1, 384), 3)       # zero question: every score is 0 -> ties -> 0,1,2
    assert res.topk_idx.cpu().tolist() == [[0, 1, 2]]
    assert (res.similarities[0].cpu().numpy() == 0).all()


def test_workspace_left_clean_and_repeatable():
    batch = synth.make_text_batch("C2", docs=16)
    a = run_case(batch["text_embeddings"], batch["question_embeddings"], 5, algo=TMA)
    b = run_case(batch["text_embeddings"], batch["question_embeddings"], 5, algo=TMA)
    c = run_case(batch["text_embeddings"], batch["question_embeddings"], 5, algo=LDG)
    f1 = run_case(batch["text_embeddings"], batch["question_embeddings"], 5, algo=CLUSTER)
    f2 = run_case(batch["text_embeddings"], batch["question_embeddings"], 5, algo=CLUSTER)
    assert torch.equal(a.topk_idx, b.topk_idx) and torch.equal(a.topk_idx, c.topk_idx)
    assert torch.equal(a.topk_idx, f1.topk_idx) and torch.equal(f1.topk_idx, f2.topk_idx)
    assert torch.equal(a.sims, b.sims)          # deterministic: fixed summation order
    assert torch.equal(a.sims, c.sims) and torch.equal(a.sims, f1.sims)   # same per-lane order and butterfly everywhere


def test_tma_rejects_oversized_tiles():
    from rag_docvqa_b200 import _lib
    batch = synth.make_text_batch("C2", docs=4)
    with pytest.raises(_lib.RdvError):
        run_case(batch["text_embeddings"], batch["question_embeddings"], 5, tile_rows=40, algo=TMA)


@pytest.mark.parametrize("algo", ALGOS)
def test_c3_slice_large_docs(algo):
    # documents above the selection cache (8192 scores) exercise the L2-resident selection path
    sizes = [20000, 13000, 500]
    emb, q = synth.make_embeddings(sizes, 768, 5, dup_frac=0.01)
    res = run_case(emb, q, 10, algo=algo)
    check_against_oracle(res, emb, q, 10)


@pytest.mark.parametrize("B", [1, 3, 97, 300])
def test_cluster_kernel_equals_two_launches(B):
    """Tiles packed into clusters (documents of 0 .. 2048 rows: padding CTAs, one tile per CTA, several tiles per CTA):
    same answers as the two-launch path, bit for bit."""
    sizes = [int(x) for x in np.random.RandomState(B).randint(0, 200, size=B)]
    sizes[0] = 70
    if B > 2:
        sizes[1], sizes[2] = 1024, 513
    emb, q = synth.make_embeddings(sizes, 256, 5 + B, dup_frac=0.1)
    a = run_case(emb, q, 7, algo=CLUSTER)
    b = run_case(emb, q, 7, algo=LDG)
    check_against_oracle(a, emb, q, 7)
    assert torch.equal(a.topk_idx, b.topk_idx) and torch.equal(a.topk_cnt, b.topk_cnt)
    assert torch.equal(a.sims, b.sims) and torch.equal(a.topk_val, b.topk_val)


def test_cluster_kernel_limits():
    from rag_docvqa_b200 import _lib
    from rag_docvqa_b200 import functional as F
    dev = torch.device("cuda:0")
    emb, q = synth.make_embeddings([3000, 10], 128, 3)
    with pytest.raises(_lib.RdvError):              # a document above rdv_cluster_max_rows(): no cluster table was built
        F.score_topk([e.to(dev) for e in emb], q.to(dev), 5, cluster=True)
    emb, q = synth.make_embeddings([100, 10], 128, 3)
    with pytest.raises(_lib.RdvError):              # k above rdv_cluster_max_k()
        F.score_topk([e.to(dev) for e in emb], q.to(dev), 40, cluster=True)
    res = F.score_topk([e.to(dev) for e in emb], q.to(dev), 40)      # the product's own choice falls back to two launches
    check_against_oracle(res, emb, q, 40)


def test_rejects_cpu_tensors():
    from rag_docvqa_b200 import functional as F
    with pytest.raises(RuntimeError):
        F.score_topk([torch.randn(4, 8)], torch.randn(1, 8), 2)


def test_c3_slice_32_documents_at_full_length():
    """32 documents of the full C3 length (10 000 chunks x 768-d each, 983 MB) against the oracle: scores, hits, order."""
    sizes = [10_000] * 30 + [9_999, 10_000]
    emb, q = synth.make_embeddings(sizes, 768, 41, dup_frac=0.01)
    res = run_case(emb, q, 10)
    check_against_oracle(res, emb, q, 10)


@pytest.mark.parametrize("algo", [0] + ALGOS)
def test_fuzz_shapes_against_the_oracle(algo):
    """Forty batches nobody chose (fixed seeds): 1-12 documents of 0-1500 rows, d in {4 .. 1024} (multiples of 4), k in 1 .. 32,
    duplicated rows (exact ties), zero rows, rows scaled by 1e-20 / 1e20, a zero question -- every path must agree with the
    oracle on the scores and, on its own scores, on the order."""
    rng = np.random.RandomState(4242)
    for case in range(40):
        B = int(rng.randint(1, 13))
        d = 4 * int(rng.choice([1, 2, 8, 24, 25, 48, 96, 160, 192, 256]))
        if algo == TMA:                                       # the TMA-ring kernel exists for these widths only (rdv_score_plan says so)
            d = int(rng.choice([128, 256, 384, 512, 768, 1024]))
        k = int(rng.randint(1, 33))
        g = torch.Generator().manual_seed(1000 + case)
        emb = []
        for b in range(B):
            n = int(rng.choice([0, 1, 2, 7, 31, 33, 64, 300, 600, 1500], p=[.08, .08, .08, .1, .1, .1, .1, .16, .1, .1]))
            e = torch.randn(n, d, generator=g)
            if n > 3:
                e[n // 2] = e[0]                                  # exact duplicate: lowest index first
                e[n - 1] = 0.0                                    # zero row: score exactly 0
                e[1] *= 1e-20
                e[2] *= 1e20
            emb.append(e)
        q = torch.randn(B, d, generator=g)
        if B > 2:
            q[2] = 0.0
        res = run_case(emb, q, k, algo=algo)
        check_against_oracle(res, emb, q, k)
