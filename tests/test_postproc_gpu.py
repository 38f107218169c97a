"""GPU parity of the reranker post-processing and the page vote: rdv_rerank_order,
rdv_page_vote and the gather kernel's emit_order against the oracle and the reference's frozen outputs."""
import zlib

import numpy as np
import pytest
import torch

from oracle import ref_restated as R
from rag_docvqa_b200 import synth
from oracle.compare import assert_order_matches_modulo_ties, load_postproc_golden as load_postproc

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
BASE = {"compute_stats": False, "compute_stats_examples": False, "n_stats_examples": 0, "device": DEV}


def run_rerank(rows, dtype, thresh, mx, mn):
    from rag_docvqa_b200 import postproc
    k = max(1, max(len(r) for r in rows))
    pad = np.zeros((len(rows), k), dtype=dtype)
    for b, r in enumerate(rows):
        pad[b, :len(r)] = r
    cnt = torch.tensor([len(r) for r in rows], dtype=torch.int32, device=DEV)
    order, kept, out_scores = postproc.rerank_order(torch.from_numpy(pad).to(DEV), cnt, thresh, mx, mn)
    order, kept, out_scores = order.cpu().numpy(), kept.cpu().numpy(), out_scores.cpu().numpy()
    res = []
    for b, r in enumerate(rows):
        o = order[b, :kept[b]].tolist()
        assert (order[b, kept[b]:] == -1).all()
        got_s, want_s = out_scores[b, :kept[b]], np.asarray(r, dtype=dtype)[o]
        assert np.array_equal(got_s, want_s, equal_nan=True)
        res.append(o)
    return res


def test_rerank_order_matches_reference_golden(golden_dir):
    g = load_postproc(golden_dir)
    for c in g["rerank"]:
        got = run_rerank([c["scores_np"]], c["scores_np"].dtype, c["thresh"], c["max"], c["min"])[0]
        assert got == R.rerank_order(c["scores_np"], c["thresh"], c["max"], c["min"])    # bit-exact vs the oracle
        assert_order_matches_modulo_ties(got, c["order"], c["scores_np"])                 # the reference modulo ties


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_rerank_order_random_batches(dtype):
    rng = np.random.RandomState(11)
    for thresh, mx, mn in [(0.4, 5, 1), (0.4, 3, 1), (0.9, 5, 3), (0.0, 64, 1), (0.5, 2, 4), (0.4, 5, 0), (0.4, 0, 0),
                           (-1.0, 7, 70)]:
        rows = []
        for b in range(97):
            n = int(rng.choice([0, 1, 2, 5, 10, 20, 33, 64]))
            kind = b % 4
            if kind == 0:
                r = rng.rand(n)
            elif kind == 1:
                r = rng.choice([0.1, 0.4, 0.55, 0.9, -0.0, 0.0], size=n)
            elif kind == 2:
                r = rng.randn(n) * 3
            else:
                r = rng.rand(n)
                r[rng.rand(n) < 0.2] = np.nan
                r[rng.rand(n) < 0.1] = np.inf
            rows.append(r.astype(dtype))
        got = run_rerank(rows, dtype, thresh, mx, mn)
        for b, r in enumerate(rows):
            assert got[b] == R.rerank_order(r, thresh, mx, mn), (thresh, mx, mn, r.tolist())


def test_threshold_is_compared_in_float64():
    # float32(0.3) = 0.30000001192... >= 0.3 in float64 (what numpy 1.x does); the float below it is not
    below = np.nextafter(np.float32(0.3), np.float32(0))
    got = run_rerank([np.array([0.3, below], dtype=np.float32)], np.float32, 0.3, 5, 0)[0]
    assert got == [0] == R.rerank_order(np.array([0.3, below], dtype=np.float32), 0.3, 5, 0)


def vote(pages, sims, weighted, legacy):
    from rag_docvqa_b200 import postproc
    return postproc.major_page_indices(pages, [torch.from_numpy(np.asarray(s, dtype=np.float32)) for s in sims],
                                       "weightmajorpage" if weighted else "majorpage", device=DEV,
                                       legacy_promotion=legacy)


def test_page_vote_matches_reference_golden(golden_dir):
    g = load_postproc(golden_dir)
    for c in g["page_vote"]:
        weighted = c["mode"] == "weightmajorpage"
        assert vote(c["pages"], c["sims"], weighted, legacy=False) == c["major"]     # made under numpy >= 2 (NEP 50)
        want = [R.page_vote(p, s, len(s), weighted, legacy_promotion=True) for p, s in zip(c["pages"], c["sims"])]
        assert vote(c["pages"], c["sims"], weighted, legacy=True) == want


@pytest.mark.parametrize("weighted", [False, True])
@pytest.mark.parametrize("legacy", [False, True])
def test_page_vote_random_vs_oracle(weighted, legacy):
    from rag_docvqa_b200 import postproc
    rng = np.random.RandomState(3 + weighted + 2 * legacy)
    pages, sims = [], []
    for b in range(300):
        n_b = int(rng.choice([0, 1, 3, 40, 600, 4000]))
        k_b = min(int(rng.choice([1, 5, 10, 20, 64])), n_b)
        span = int(rng.choice([2, 5, 33, 200, 5000]))
        pages.append([int(x) for x in rng.randint(0, span, size=k_b)])
        s = rng.rand(n_b).astype(np.float32) - (0.3 if b % 3 == 0 else 0.0)
        if b % 7 == 0 and n_b:
            s[:] = np.float32(0.25)                        # identical weights: page ties decided by CPython's set order
        if b % 50 == 1 and n_b:
            s[0] = np.nan
        if b % 11 == 3 and n_b > 2:
            s[1] = np.float32(1e-30)                       # wide exponent span: the order of the float64 sum matters again
            s[2] = np.float32(-1e-41)                      # subnormal
        sims.append(s)
    got = vote(pages, sims, weighted, legacy)
    want = [R.page_vote(p, s, len(s), weighted, legacy_promotion=legacy) for p, s in zip(pages, sims)]
    assert got == want
    # winning weight, bit for bit
    row_off = np.zeros(len(sims) + 1, dtype=np.int64); np.cumsum([len(s) for s in sims], out=row_off[1:])
    k = 64
    hp = np.zeros((len(pages), k), dtype=np.int32)
    for b, p in enumerate(pages):
        hp[b, :len(p)] = p
    major, weight = postproc.page_vote(torch.from_numpy(hp).to(DEV),
                                       torch.tensor([len(p) for p in pages], dtype=torch.int32, device=DEV),
                                       torch.from_numpy(np.concatenate(sims)).to(DEV), torch.from_numpy(row_off).to(DEV),
                                       weighted, legacy, return_weight=True)
    assert major.cpu().tolist() == want and weight.dtype == torch.float64


def prompts_for(questions):
    return [[5 + (zlib.crc32(t.encode()) % 1000) for t in ("question: {:s}  context: ".format(qs)).split()]
            for qs in questions]


class FixedCrossEncoder:
    """Stands in for the model: score of a (question, candidate) pair = a hash of the text."""
    def __init__(self, as_list=False):
        self.as_list = as_list

    def forward(self, pairs):
        s = np.array([(zlib.crc32((q + "|" + c).encode()) % 1000) / 1000.0 for q, c in pairs], dtype=np.float32)
        return [float(x) for x in s] if self.as_list else s


@pytest.mark.parametrize("as_list", [False, True])
def test_reranker_class_matches_oracle_lists(as_list):
    from rag_docvqa_b200.postproc import Reranker
    ce = FixedCrossEncoder(as_list)
    rr = Reranker({"rerank_filter_tresh": 0.4, "rerank_max_chunk_num": 3, "rerank_min_chunk_num": 2, "device": DEV}, ce)
    questions = ["q%d" % b for b in range(9)]
    cands = [["c%d_%d" % (b, i) for i in range([0, 1, 2, 5, 5, 10, 10, 20, 7][b])] for b in range(9)]
    extra = [[(b, i) for i in range(len(c))] for b, c in enumerate(cands)]
    out_c, out_e = rr.batch_rerank(questions, cands, extra)
    for b in range(9):
        scores = ce.forward([(questions[b], c) for c in cands[b]])
        want_c, want_e = R.rerank(np.asarray(scores), cands[b], extra[b], filter_thresh=0.4, max_chunk_num=3, min_chunk_num=2)
        assert out_c[b] == want_c and out_e[b] == want_e
        one_c, one_e = rr.rerank(questions[b], cands[b], extra[b])
        assert one_c == want_c and one_e == want_e
    with pytest.raises(ValueError):
        Reranker({}, None)


@pytest.mark.parametrize("s,reorder,sep", [(0, False, False), (0, True, True), (3, True, True)])
def test_packed_rerank_equals_retrieve_then_rerank_then_pack(s, reorder, sep):
    """retrieve -> Reranker.batch_rerank -> flatten -> prepare_inputs_for_vqa (the reference's sequence,
    src/RAGVT5.py:244-282) against retrieve_packed -> rerank_packed on the device."""
    from rag_docvqa_b200.docstore import DocStore
    from rag_docvqa_b200.postproc import Reranker
    from rag_docvqa_b200.retriever import Retriever
    batch = synth.make_text_batch("C2", with_lists=True, docs=12, seed=56, dup_frac=0.0)
    words, boxes, labels = batch["words_text_chunks"], batch["words_box_chunks"], batch["layout_labels_chunks"]
    pages, images = batch["page_indices"], batch["images"]
    table = synth.make_tokens_for_words(words, seed=9)
    tok = lambda w: table.get(w, [2])
    store = DocStore.from_lists(words, boxes, labels, pages, tok, torch.device(DEV), images=images)
    questions = ["what is item %d about ?" % b for b in range(len(words))]
    prompts = prompts_for(questions)
    sep_ids = [2, 9] if sep else []
    k = 6
    retr = Retriever({**BASE, "chunk_num": k, "include_surroundings": s, "reorder_chunks": reorder})
    packed, res, plan = retr.retrieve_packed([e.to(DEV) for e in batch["text_embeddings"]],
                                             batch["question_embeddings"].to(DEV), store, prompts, sep_ids=sep_ids,
                                             max_source_length=512, with_layout_labels=True, return_plan=True)
    hits = Retriever._hits_to_host(res.topk_idx, res.topk_cnt)
    ref = R.gather_hits(hits, words, boxes, labels, images, pages, include_surroundings=s, reorder_chunks=reorder,
                        crop=False)
    before = packed.hit_chunk.cpu().numpy().copy()
    rng = np.random.RandomState(8)
    scores = rng.rand(len(words), k).astype(np.float32)
    scores[3] = 0.0                                             # nothing passes: the min fallback
    rr = Reranker({"rerank_filter_tresh": 0.4, "rerank_max_chunk_num": 4, "rerank_min_chunk_num": 1, "device": DEV},
                  FixedCrossEncoder())
    packed2, order, kept = rr.rerank_packed(plan, torch.from_numpy(scores).to(DEV))
    order, kept = order.cpu().numpy(), kept.cpu().numpy()
    new_lists = [[], [], [], [], []]      # boxes, labels, words, word boxes, word labels + pages
    new_pages, new_rects = [], []
    for b in range(len(words)):
        n_b = len(ref[0][b])
        o = R.rerank_order(scores[b, :n_b], 0.4, 4, 1)
        assert order[b, :kept[b]].tolist() == o
        for dst, src in zip(new_lists, (ref[1], ref[2], ref[3], ref[4], ref[5])):
            dst.append([src[b][i] for i in o])
        new_pages.append([ref[7][b][i] for i in o]); new_rects.append([ref[6][b][i] for i in o])
        assert packed2.hit_chunk[b, :kept[b]].cpu().tolist() == [int(before[b, i]) for i in o]
    sep_word = "<sep>" if sep else None
    tok_ref = lambda w: sep_ids if w == "<sep>" else tok(w)
    ids, bxs, mask, labs = R.vt5_pack(prompts, [R.flatten(x, sep_word) for x in new_lists[2]],
                                      [R.flatten(x, sep_word) for x in new_lists[3]], tok_ref,
                                      layout_labels=[R.flatten(x, sep_word) for x in new_lists[4]])
    assert torch.equal(packed2.input_ids.cpu(), ids)
    assert torch.equal(packed2.boxes.cpu(), bxs)
    assert torch.equal(packed2.attention_mask.cpu(), mask)
    assert torch.equal(packed2.layout_labels.cpu(), labs)
    bbox = packed2.hit_bbox.cpu().numpy(); rect = packed2.hit_rect.cpu().numpy(); page = packed2.hit_page.cpu().numpy()
    for b in range(len(words)):
        for j in range(kept[b]):
            assert bbox[b, j].tolist() == [float(x) for x in new_lists[0][b][j]]
            assert rect[b, j].tolist() == new_rects[b][j] and page[b, j] == new_pages[b][j]
        assert (page[b, kept[b]:] == -1).all()
    # the vote over the reranked pages, straight from the device arrays
    from rag_docvqa_b200 import postproc
    row_off = torch.from_numpy(np.concatenate([[0], np.cumsum(res.sizes)]).astype(np.int64)).to(DEV)
    major = postproc.page_vote(packed2.hit_page, kept if isinstance(kept, torch.Tensor) else torch.from_numpy(kept).to(DEV),
                               res.sims, row_off, weighted=True)
    sims_h = [x.cpu().numpy() for x in res.similarities]
    assert major.cpu().tolist() == [R.page_vote(new_pages[b], sims_h[b], len(sims_h[b]), True) for b in range(len(words))]
