"""GPU checks of the bf16 tensor-core modes (tcgen05 / TMEM / TMA): corpus score + fused top-k, MaxSim.

bf16 mode is not a bit-parity mode (north_star): it is held to (a) the exact result of the SAME bf16
operands computed in float64 on the CPU, within fp32-accumulation error, and (b) recall@k against the
fp32 oracle on the un-rounded inputs."""
import numpy as np
import pytest
import torch

from oracle import compare
from oracle import ref_restated as R

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(params=["1", "2"], autouse=True, ids=["one_cta", "cta_pair"])
def tc_ctas(request, monkeypatch):
    """Every test runs with the single-CTA kernel and with CTA pairs (tcgen05 cta_group::2; used when the
    question tiles pair up).  librdv reads RDV_TC_CTAS at each launch."""
    monkeypatch.setenv("RDV_TC_CTAS", request.param)


def bf16_round(x):
    return x.to(torch.bfloat16).to(torch.float64)


def exact_scores_of_bf16_operands(E, Q):
    Eb, Qb = bf16_round(E), bf16_round(Q)
    dots = Qb @ Eb.T
    return (dots / (Qb.norm(dim=-1)[:, None] * Eb.norm(dim=-1)[None, :])).numpy()


@pytest.mark.parametrize("n,d,Qn,k", [(256, 64, 128, 5), (1000, 128, 37, 10), (5000, 768, 300, 10), (70000, 384, 130, 16),
                                      (300, 72, 5, 3), (20000, 768, 512, 10), (3000, 256, 1000, 4)])
def test_corpus_topk_matches_exact_bf16_math(n, d, Qn, k, check_recall=True):
    from rag_docvqa_b200.sharded import CorpusShard
    g = torch.Generator().manual_seed(n + d)
    u = torch.randn(d, generator=g)
    E = torch.randn(n, d, generator=g) + 0.5 * u
    Q = torch.randn(Qn, d, generator=g) + 0.5 * u
    E[7] = E[3]                                       # exact duplicate rows: lowest id first
    shard = CorpusShard.from_f32(E.to(DEV), id_offset=1000)
    val, idx = shard.search_local(Q.to(DEV), k)
    torch.cuda.synchronize()
    val, idx = val.cpu().numpy(), idx.cpu().numpy()
    exact = exact_scores_of_bf16_operands(E, Q)
    assert idx.shape == (Qn, k)
    for q in range(Qn):
        ids = idx[q] - 1000
        assert (ids >= 0).all() and (ids < n).all() and len(set(ids.tolist())) == k
        np.testing.assert_allclose(val[q], exact[q][ids], rtol=2e-5, atol=2e-6)
        # the selection is the exact top-k of those scores up to fp32 accumulation noise at the boundary
        kth = np.sort(exact[q])[::-1][k - 1]
        assert (exact[q][ids] >= kth - 1e-5).all()
        assert set(np.nonzero(exact[q] > kth + 1e-5)[0].tolist()) <= set(ids.tolist())
        assert (np.diff(val[q]) <= 0).all()
        if 3 in ids and 7 in ids:
            assert list(ids).index(3) < list(ids).index(7)
    # recall@k against the fp32 oracle on the un-rounded inputs
    if not check_recall:
        return
    ref = R.corpus_scores(E, Q).numpy()
    ref_idx = np.stack([R.topk_lowest_index(ref[q], k) for q in range(Qn)])
    assert compare.recall_at_k(idx - 1000, ref_idx) >= 0.9


def test_corpus_sharded_equals_unsharded():
    from rag_docvqa_b200 import sharded
    g = torch.Generator().manual_seed(5)
    n, d, Qn, k, world = 9000, 256, 200, 10, 4
    E = torch.randn(n, d, generator=g)
    Q = torch.randn(Qn, d, generator=g)
    E[8000] = E[10]                                   # a cross-shard exact tie
    full = sharded.CorpusShard.from_f32(E.to(DEV))
    v_full, i_full = full.search_local(Q.to(DEV), k)
    vals, idxs = [], []
    for r in range(world):                            # ranks emulated as slices on one GPU (B200 guide)
        lo, hi = sharded.shard_bounds(n, world, r)
        sh = sharded.CorpusShard.from_f32(E[lo:hi].to(DEV), id_offset=lo)
        v, i = sh.search_local(Q.to(DEV), k)
        vals.append(v); idxs.append(i)
    v_m, i_m = sharded.merge_candidates(torch.cat(vals, 1), torch.cat(idxs, 1), k)
    assert torch.equal(i_m, i_full)
    assert torch.equal(v_m, v_full)


@pytest.mark.parametrize("n,Lq,Lp,d", [(3, 128, 256, 64), (5, 200, 300, 128), (4, 2048, 2048, 768), (2, 77, 513, 96)])
def test_maxsim_bf16_tc(n, Lq, Lp, d):
    from rag_docvqa_b200 import functional as F
    g = torch.Generator().manual_seed(n + Lq)
    q = torch.randn(1, Lq, d, generator=g)
    p = torch.randn(n, Lp, d, generator=g)
    got = F.late_interaction_bf16(q.to(DEV), p.to(DEV)).cpu().numpy()
    # exact math on the same normalised-then-rounded operands
    qn = bf16_round(torch.nn.functional.normalize(q, dim=-1))
    pn = bf16_round(torch.nn.functional.normalize(p, dim=-1))
    exact = torch.bmm(qn.expand(n, -1, -1), pn.transpose(1, 2)).max(dim=-1).values.sum(dim=-1).numpy()
    np.testing.assert_allclose(got, exact, rtol=2e-5)
    ref = R.late_interaction_f64(q, p).numpy()          # fp32-mode truth: bf16 rounding error only
    np.testing.assert_allclose(got, ref, rtol=5e-3)
    fp32 = F.late_interaction(q.to(DEV), p.to(DEV)).cpu().numpy()
    assert (np.argsort(-got)[:1] == np.argsort(-fp32)[:1]).all() or n < 2


def test_corpus_searcher_graph_equals_search_local():
    """CorpusSearcher (static buffers, the step captured into one CUDA graph) returns exactly what the eager calls do,
    call after call, and its merge over the receive layout (rdv_topk_merge_parts) equals the unsharded answer."""
    from rag_docvqa_b200 import _lib, sharded
    g = torch.Generator().manual_seed(9)
    n, d, Qn, k, world = 9000, 256, 130, 10, 4
    E = torch.randn(n, d, generator=g)
    E[8000] = E[10]                                   # a cross-shard exact tie
    full = sharded.CorpusShard.from_f32(E.to(DEV))
    searcher = sharded.CorpusSearcher(full, Qn, k, graph=True)
    assert searcher.graphed
    for seed in (1, 2, 3):
        Q = torch.randn(Qn, d, generator=torch.Generator().manual_seed(seed)).to(DEV)
        v_ref, i_ref = full.search_local(Q, k)
        v, i = searcher.search(Q)
        assert torch.equal(i, i_ref) and torch.equal(v, v_ref)
    # ranks emulated as slices on one GPU: each rank's searcher fills its send buffer; the receive buffer is their
    # concatenation in rank order, exactly what ncclAllGather writes
    sends = []
    for r in range(world):
        lo, hi = sharded.shard_bounds(n, world, r)
        s_r = sharded.CorpusSearcher(sharded.CorpusShard.from_f32(E[lo:hi].to(DEV), id_offset=lo), Qn, k, graph=(r % 2 == 0))
        s_r.local_only(Q)
        sends.append(s_r.send.clone())
        nv, stride = s_r.nv, s_r.send.numel()
    recv = torch.stack(sends).contiguous()
    out_v = torch.empty((Qn, k), dtype=torch.float32, device=DEV)
    out_i = torch.empty((Qn, k), dtype=torch.int64, device=DEV)
    _lib.check(_lib.lib.rdv_topk_merge_parts(recv.data_ptr(), recv.data_ptr() + nv, Qn, world, k, stride // 4, stride // 8, k,
                                             out_v.data_ptr(), out_i.data_ptr(), None))
    torch.cuda.synchronize()
    assert torch.equal(out_i, i_ref) and torch.equal(out_v, v_ref)


def test_corpus_searcher_empty_shard():
    """A rank that owns no rows (N < world) still takes part: its contribution is 'no candidates'."""
    from rag_docvqa_b200 import sharded
    empty = sharded.CorpusShard(torch.empty((0, 64), dtype=torch.bfloat16, device=DEV), id_offset=5)
    s = sharded.CorpusSearcher(empty, 7, 3, graph=True)
    v, i = s.search(torch.randn(7, 64, device=DEV))
    assert (i == -1).all() and torch.isneginf(v).all()
    v2, i2 = empty.search_local(torch.randn(7, 64, device=DEV), 3)
    assert (i2 == -1).all() and torch.isneginf(v2).all()


def test_two_devices_in_one_process():
    """Kernel attributes (dynamic shared memory above 48 KB, cluster size) belong to a device's context: a process that
    uses cuda:0 and then cuda:1 must be able to launch the tcgen05, TMA-ring, pooling, gather and cluster kernels on both
    (the opt-in used to be remembered once per process).  Needs >= 2 GPUs."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs (gpurun --gpus 2)")
    from rag_docvqa_b200 import functional as F
    from rag_docvqa_b200 import sharded, synth
    results = []
    for dev in ("cuda:0", "cuda:1", "cuda:0"):
        g = torch.Generator().manual_seed(3)
        E = torch.randn(3000, 256, generator=g)
        Q = torch.randn(130, 256, generator=g)
        with torch.cuda.device(dev):
            v, i = sharded.CorpusShard.from_f32(E.to(dev)).search_local(Q.to(dev), 10)            # tc_score_kernel (~209 KB)
            emb, q = synth.make_embeddings([300, 17, 0, 600], 384, 5)
            a = F.score_topk([e.to(dev) for e in emb], q.to(dev), 5, algo=2)                      # score_tma_kernel (192 KB)
            b = F.score_topk([e.to(dev) for e in emb], q.to(dev), 5, cluster=True)                # clusters of 16
            qq, pp = torch.randn(1, 256, 128, generator=g), torch.randn(3, 256, 128, generator=g)
            m = F.late_interaction(qq.to(dev), pp.to(dev), mode="tf32x3")                         # maxsim_tf32x3_kernel (192 KB)
            embs, am = synth.make_token_batch(64, 1024, 7, max_len=40)
            p = F.mean_pooling(embs.to(dev), am.to(dev), normalise=True)                          # mean_pool_kernel
            torch.cuda.synchronize(dev)
        assert torch.equal(a.topk_idx.cpu(), b.topk_idx.cpu())
        results.append((i.cpu(), a.topk_idx.cpu(), m.cpu(), p.cpu()))
    for r in results[1:]:
        for x, y in zip(results[0], r):
            assert torch.equal(x, y)


@pytest.mark.parametrize("ctas", ["1", "2", ""])
def test_corpus_topk_fuzz(monkeypatch, ctas):
    """A dozen seeded (rows, width, questions, k) combinations nobody chose -- partial row tiles, partial question tiles,
    widths that are not a multiple of the 64-wide k-block, k from 1 to 16 -- through the single-CTA kernel, the CTA-pair
    kernel and the launcher's own choice."""
    if ctas:
        monkeypatch.setenv("RDV_TC_CTAS", ctas)
    else:
        monkeypatch.delenv("RDV_TC_CTAS", raising=False)
    rng = np.random.RandomState(7 + len(ctas))
    for _ in range(12):
        n = int(rng.choice([257, 511, 1024, 3333, 9000, 40001]))
        d = 8 * int(rng.choice([1, 3, 8, 9, 16, 48, 96, 128]))
        Qn = int(rng.choice([1, 2, 127, 128, 129, 256, 300, 640]))
        k = int(rng.randint(1, 17))
        # recall against fp32 is a statistic: only where the sample is large and the rows are not near-ties by construction
        test_corpus_topk_matches_exact_bf16_math(n, d, Qn, min(k, n), check_recall=(Qn * k >= 500 and d >= 128))
