"""GPU parity of the drop-in classes against the reference's frozen outputs (tests/golden) and the
oracle: Retriever.retrieve 9-tuple, VisualRetriever.retrieve, and the device gather (packed VT5 inputs)."""
import json
import os
import zlib

import numpy as np
import pytest
import torch

from oracle import compare
from oracle import ref_restated as R
from rag_docvqa_b200 import synth

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
BASE = {"compute_stats": False, "compute_stats_examples": False, "n_stats_examples": 0, "device": DEV}


def digest(im):
    return [im.size[0], im.size[1], zlib.crc32(im.convert("RGB").tobytes()) & 0xFFFFFFFF]


def load_retrieve_inputs(golden_dir):
    with open(os.path.join(golden_dir, "retrieve_lists.json")) as f:
        gold = json.load(f)
    z = np.load(os.path.join(golden_dir, "retrieve_inputs.npz"))
    sizes, cpp = gold["sizes"], gold["chunks_per_page"]
    emb = [torch.from_numpy(z["emb_%d" % b]) for b in range(len(sizes))]
    q = torch.from_numpy(z["q"])
    words, boxes, labels = synth.make_words(sizes, gold["words_seed"], min_words=3, max_words=9, empty_chunk_every=11)
    pages = synth.make_page_indices(sizes, cpp)
    images = synth.make_images(sizes, cpp, width=gold["image_wh"][0], height=gold["image_wh"][1], ragged_sizes=True)
    return gold, emb, q, words, boxes, labels, images, pages


def assert_same_lists(out, ref):
    for i in (0, 1, 2, 3, 4, 5, 7):
        assert out[i] == ref[i], "output %d differs" % i
    assert [[digest(im) for im in doc] for doc in out[6]] == ref[6]


@pytest.mark.parametrize("on_device", [True, False, "pinned", "host_general", "pinned_general"])
def test_retrieve_matches_reference_golden(golden_dir, on_device, monkeypatch):
    """Device tensors (the reference's case), host tensors through the small-batch path (one blob up, one launch, one
    read-back) and, with that path switched off, pageable host tensors (packed upload) and pinned host tensors
    (zero-copy: the score kernel reads the page-locked rows over PCIe) must give the same 9-tuple."""
    from rag_docvqa_b200 import retriever as retriever_module
    from rag_docvqa_b200.retriever import Retriever
    if isinstance(on_device, str) and on_device.endswith("_general"):
        monkeypatch.setattr(retriever_module, "_SMALL_BATCH_BYTES", -1)
        on_device = "pinned" if on_device.startswith("pinned") else False
    gold, emb, q, words, boxes, labels, images, pages = load_retrieve_inputs(golden_dir)
    if on_device == "pinned":
        emb, q, on_device = [e.pin_memory() for e in emb], q.pin_memory(), False
    emb_in = [e.to(DEV) for e in emb] if on_device else emb
    q_in = q.to(DEV) if on_device else q
    for var in gold["variants"]:
        retr = Retriever({**BASE, "chunk_num": var["k"], "include_surroundings": var["include_surroundings"],
                          "reorder_chunks": var["reorder_chunks"]})
        out = retr.retrieve(emb_in, q_in, words, boxes, labels, images, pages)
        ref = [var[key] for key in ("top_k_text", "top_k_boxes", "top_k_layout_labels", "top_k_words_text",
                                    "top_k_words_boxes", "top_k_words_layout_labels", "top_k_patches",
                                    "top_k_page_indices")]
        assert_same_lists(out, ref[:6] + [ref[6]] + [ref[7]])
        assert len(out[8]) == len(emb)
        for b, s in enumerate(out[8]):
            assert s.is_cuda == on_device
            compare.assert_scores_close(s.cpu().numpy(), np.array(var["similarities"][b], dtype=np.float32))


@pytest.mark.parametrize("where", ["device", "pinned_host"])
@pytest.mark.parametrize("s,reorder", [(0, False), (0, True), (2, False), (5, True), (40, False)])
def test_retrieve_c2_slice_vs_oracle(s, reorder, where):
    from rag_docvqa_b200.retriever import Retriever
    batch = synth.make_text_batch("C2", with_lists=True, docs=40 if where == "pinned_host" else 10, seed=77, dup_frac=0.0)
    args = (batch["words_text_chunks"], batch["words_box_chunks"], batch["layout_labels_chunks"],
            batch["images"], batch["page_indices"])
    retr = Retriever({**BASE, "chunk_num": 5, "include_surroundings": s, "reorder_chunks": reorder})
    if where == "device":
        out = retr.retrieve([e.to(DEV) for e in batch["text_embeddings"]], batch["question_embeddings"].to(DEV), *args)
    else:     # >= 32 documents on the host: the pipelined path (4 groups), zero-copy reads of the pinned rows
        out = retr.retrieve([e.pin_memory() for e in batch["text_embeddings"]], batch["question_embeddings"].pin_memory(), *args)
        assert not out[8][0].is_cuda and len(out[8]) == 40
    # the oracle gathers for the SAME hits (index parity is covered in test_score_topk_gpu.py)
    sims = [x.cpu() for x in out[8]]
    hits = [R.topk_lowest_index(x, 5) for x in sims]
    ref_sims = R.score(batch["text_embeddings"], batch["question_embeddings"])
    for b in range(len(hits)):
        compare.assert_topk_matches(hits[b], sims[b].numpy(), ref_sims[b].numpy(), 5)
    ref = R.gather_hits(hits, *args, include_surroundings=s, reorder_chunks=reorder)
    for i in (0, 1, 2, 3, 4, 5, 7):
        assert out[i] == ref[i], "output %d differs" % i
    assert [[digest(im) for im in doc] for doc in out[6]] == [[digest(im) for im in doc] for doc in ref[6]]


SMALL_SIZES = {1: [30], 3: [45, 0, 3], 7: [60, 0, 3, 150, 1, 90, 31]}


@pytest.mark.parametrize("pinned", [False, True])
@pytest.mark.parametrize("docs,k,s,reorder", [(1, 5, 0, False), (1, 40, 0, False), (3, 5, 0, True), (7, 5, 3, False),
                                              (7, 1, 0, False)])
def test_retrieve_small_host_batch(docs, k, s, reorder, pinned, monkeypatch):
    """Host batches of < 8 documents and <= 1 MB (C1, the reference's CPU-runnable case) take _retrieve_host_small: checked
    against the oracle and against the general host path on the same inputs; duplicates (ties), an empty document,
    documents with fewer than k chunks and a strided embedding matrix included."""
    from rag_docvqa_b200 import retriever as retriever_module
    from rag_docvqa_b200.retriever import Retriever
    sizes, dim, cpp = SMALL_SIZES[docs], 384, 30
    emb, q = synth.make_embeddings(sizes, dim, 311 + docs, dup_frac=0.1)
    words, boxes, labels = synth.make_words(sizes, 318 + docs, min_words=3, max_words=20, empty_chunk_every=13)
    lists = (words, boxes, labels, synth.make_images(sizes, cpp, width=212, height=275, ragged_sizes=True),
             synth.make_page_indices(sizes, cpp))
    wide = torch.randn(sizes[0], 2 * dim, generator=torch.Generator().manual_seed(5))
    wide[:, ::2] = emb[0]
    emb[0] = wide[:, ::2]                           # non-contiguous rows
    assert not emb[0].is_contiguous()
    assert (sum(sizes) + docs) * dim * 4 <= retriever_module._SMALL_BATCH_BYTES
    emb_in = [e.pin_memory() if pinned and e.is_contiguous() else e for e in emb]
    retr = Retriever({**BASE, "chunk_num": k, "include_surroundings": s, "reorder_chunks": reorder})
    taken = []
    small = Retriever._retrieve_host_small
    monkeypatch.setattr(Retriever, "_retrieve_host_small", lambda self, *a: (taken.append(1), small(self, *a))[1])
    for rep in range(2):                            # the second call reuses the buffers of the first
        out = retr.retrieve(emb_in, q.pin_memory() if pinned else q, *lists)
    assert len(taken) == 2
    sims = out[8]
    assert len(sims) == docs and all(not x.is_cuda and x.dtype == torch.float32 for x in sims)
    assert [x.shape[0] for x in sims] == sizes
    ref_sims = R.score([e.contiguous() for e in emb], q)
    hits = [R.topk_lowest_index(x, k) for x in sims]
    for b in range(docs):
        compare.assert_scores_close(sims[b].numpy(), ref_sims[b].numpy())
        compare.assert_topk_matches(hits[b], sims[b].numpy(), ref_sims[b].numpy(), k)
    ref = R.gather_hits(hits, *lists, include_surroundings=s, reorder_chunks=reorder)
    for i in (0, 1, 2, 3, 4, 5, 7):
        assert out[i] == ref[i], "output %d differs" % i
    assert [[digest(im) for im in doc] for doc in out[6]] == [[digest(im) for im in doc] for doc in ref[6]]
    # the general host path (packed upload, two launches) on the same inputs
    monkeypatch.setattr(retriever_module, "_SMALL_BATCH_BYTES", -1)
    gen = retr.retrieve(emb_in, q, *lists)
    assert len(taken) == 2
    for i in (0, 1, 2, 3, 4, 5, 7):
        assert out[i] == gen[i], "output %d differs from the general path" % i
    for b in range(docs):
        np.testing.assert_allclose(sims[b].numpy(), gen[8][b].numpy(), rtol=1e-6, atol=1e-7)


def test_retriever_contract_attributes():
    from rag_docvqa_b200.retriever import Retriever, VisualRetriever
    with pytest.raises(KeyError):
        Retriever({"chunk_num": 3})                      # the three stat keys are required (reference :180-182)
    r = Retriever({**BASE, "compute_stats": True, "layout_model": "YOLO"})
    assert r.k == 10 and r.include_surroundings == 0 and r.reorder_chunks is False
    assert r.stats == {"layout_labels_topk_dist": {"title": 0, "text": 0, "figure": 0, "table": 0}}
    assert r.layout_map[1] == "text"
    v = VisualRetriever({"chunk_num": 5, "device": DEV})
    assert v.k == 5 and v.mode == "horizontal"
    out = Retriever({**BASE, "chunk_num": 3}).retrieve([torch.zeros(0, 8)], torch.zeros(1, 8).to(DEV), [[]], [[]], [[]], [[]], [[]])
    assert [o for o in out[:8]] == [[[]]] * 8 and out[8][0].shape == (0,)


def test_visual_retrieve_golden(golden_dir):
    from PIL import Image
    from rag_docvqa_b200.retriever import VisualRetriever
    with open(os.path.join(golden_dir, "visual_retrieve.json")) as f:
        gold = json.load(f)
    z = np.load(os.path.join(golden_dir, "visual_inputs.npz"))
    rng = np.random.RandomState(5)
    flat, mats, xyxy, images, patches = [], [], [], [], []
    for b, doc in enumerate(gold["docs"]):
        f_b, imgs, m_b = [], [], []
        for g, n_rows in enumerate(doc["groups"]):
            W, H = doc["image_wh"][g]
            page = Image.fromarray(rng.randint(0, 255, size=(H, W, 3)).astype(np.uint8), "RGB")
            imgs.append(page)
            m_b.append([[page.crop(tuple(rc))] for rc in doc["xyxy"][g]])
            f_b.extend([g] * n_rows)
        flat.append(np.array(f_b, dtype=np.int64)); mats.append(m_b); xyxy.append(doc["xyxy"]); images.append(imgs)
        patches.append(torch.from_numpy(z["p_%d" % b]).to(DEV))
    q = torch.from_numpy(z["q"]).to(DEV)
    for var in gold["variants"]:
        s = tuple(var["include_surroundings"]) if isinstance(var["include_surroundings"], list) else var["include_surroundings"]
        vr = VisualRetriever({"chunk_num": var["k"], "include_surroundings": s, "chunk_mode": "horizontal", "device": DEV})
        sims = vr._get_similarities(patches, q)
        for b in range(len(patches)):
            np.testing.assert_allclose(sims[b].cpu().numpy(), np.array(var["sims"][b], dtype=np.float32), rtol=1e-5)
        crops, page_ids = vr.retrieve(patches, q, flat, mats, xyxy, images)
        assert [sorted(digest(im) for im in doc) for doc in crops] == var["crops"]
        assert [sorted(doc) for doc in page_ids] == var["pages"]
    with pytest.raises(NotImplementedError):
        VisualRetriever({"chunk_num": 1, "chunk_mode": "square", "device": DEV}).retrieve(patches, q, flat, mats, xyxy, images)


# ---------------------------------------------------------------------------------------------------
# device gather -> packed VT5 inputs
# ---------------------------------------------------------------------------------------------------
def prompts_for(questions):
    return [[5 + (zlib.crc32(t.encode()) % 1000) for t in ("question: {:s}  context: ".format(qs)).split()]
            for qs in questions]


@pytest.mark.parametrize("derived", [True, False])
def test_packed_inputs_match_reference_golden(golden_dir, derived):
    """tests/golden/vt5_pack.json holds what the reference's flatten + VT5.prepare_inputs_for_vqa built."""
    from rag_docvqa_b200.docstore import DocStore
    from rag_docvqa_b200.retriever import Retriever
    gold, emb, q, words, boxes, labels, images, pages = load_retrieve_inputs(golden_dir)
    with open(os.path.join(golden_dir, "vt5_pack.json")) as f:
        packs = json.load(f)
    table = synth.make_tokens_for_words(words, seed=packs["word_table_seed"])
    store = DocStore.from_lists(words, boxes, labels, pages, lambda w: table.get(w, [2]), torch.device(DEV), images=images,
                                derived=derived)
    prompts = prompts_for(packs["questions"])
    retr = Retriever({**BASE, "chunk_num": packs["k"]})
    for var in packs["variants"]:
        sep_ids = [2] if var["sep"] else []                      # FakeTokenizer maps unknown words to [2]
        packed, res = retr.retrieve_packed([e.to(DEV) for e in emb], q.to(DEV), store, prompts, sep_ids=sep_ids,
                                           max_source_length=var["max_source_length"],
                                           with_layout_labels=var["use_layout_labels"] == "Embed")
        assert packed.input_ids.cpu().tolist() == var["input_ids"]
        assert packed.boxes.cpu().tolist() == var["boxes"]
        assert packed.attention_mask.cpu().tolist() == var["attention_mask"]
        if var["layout_labels"] is not None:
            assert packed.layout_labels.cpu().tolist() == var["layout_labels"]


@pytest.mark.parametrize("derived", [True, False])
@pytest.mark.parametrize("s,reorder,sep", [(0, False, False), (0, True, True), (3, False, True), (7, True, True),
                                           (200, True, False)])
def test_packed_inputs_vs_oracle_c2_slice(s, reorder, sep, derived):
    from rag_docvqa_b200.docstore import DocStore
    from rag_docvqa_b200.retriever import Retriever
    batch = synth.make_text_batch("C2", with_lists=True, docs=12, seed=55, dup_frac=0.0)
    words, boxes, labels = batch["words_text_chunks"], batch["words_box_chunks"], batch["layout_labels_chunks"]
    pages, images = batch["page_indices"], batch["images"]
    table = synth.make_tokens_for_words(words, seed=9)
    tok = lambda w: table.get(w, [2])
    store = DocStore.from_lists(words, boxes, labels, pages, tok, torch.device(DEV), images=images, derived=derived)
    questions = ["what is item %d about ?" % b for b in range(len(words))]
    prompts = prompts_for(questions)
    sep_ids = [2, 9] if sep else []
    retr = Retriever({**BASE, "chunk_num": 5, "include_surroundings": s, "reorder_chunks": reorder})
    packed, res = retr.retrieve_packed([e.to(DEV) for e in batch["text_embeddings"]],
                                       batch["question_embeddings"].to(DEV), store, prompts, sep_ids=sep_ids,
                                       max_source_length=512, with_layout_labels=True, )
    hits = Retriever._hits_to_host(res.topk_idx, res.topk_cnt)
    ref = R.gather_hits(hits, words, boxes, labels, images, pages, include_surroundings=s, reorder_chunks=reorder,
                        crop=False)
    sep_word = "<sep>" if sep else None
    tok_ref = lambda w: sep_ids if w == "<sep>" else tok(w)
    ids, bxs, mask, labs = R.vt5_pack(prompts, [R.flatten(x, sep_word) for x in ref[3]],
                                      [R.flatten(x, sep_word) for x in ref[4]], tok_ref,
                                      layout_labels=[R.flatten(x, sep_word) for x in ref[5]])
    assert torch.equal(packed.input_ids.cpu(), ids)
    assert torch.equal(packed.boxes.cpu(), bxs)
    assert torch.equal(packed.attention_mask.cpu(), mask)
    assert torch.equal(packed.layout_labels.cpu(), labs)
    # hit metadata: bbox (a8), crop rectangle (a9), page / label, in output order
    bbox = packed.hit_bbox.cpu().numpy(); rect = packed.hit_rect.cpu().numpy()
    page = packed.hit_page.cpu().numpy(); label = packed.hit_label.cpu().numpy()
    for b in range(len(words)):
        for j in range(len(ref[1][b])):
            assert bbox[b, j].tolist() == [float(x) for x in ref[1][b][j]]
            assert rect[b, j].tolist() == ref[6][b][j]
            assert page[b, j] == ref[7][b][j] and label[b, j] == ref[2][b][j]
        assert (page[b, len(ref[1][b]):] == -1).all()


@pytest.mark.parametrize("derived", [True, False])
@pytest.mark.parametrize("reorder,sep,k", [(False, False, 5), (True, True, 5), (True, False, 20)])
def test_one_launch_step_equals_two_launches(reorder, sep, k, derived):
    """rdv_retrieve_vt5_f32 (score + top-k + gather in ONE launch, a thread-block cluster per document, the candidates'
    chunk records handed over through distributed shared memory) against rdv_score_f32 followed by rdv_gather_vt5_inputs
    with its own selection: every output bit for bit, twice."""
    from rag_docvqa_b200 import functional as F
    from rag_docvqa_b200.docstore import DocStore
    batch = synth.make_text_batch("C2", with_lists=True, docs=24, seed=77, dup_frac=0.1)
    words, boxes, labels = batch["words_text_chunks"], batch["words_box_chunks"], batch["layout_labels_chunks"]
    sizes = batch["sizes"]
    assert 0 in sizes and any(0 < n < 5 for n in sizes) and max(sizes) > 256       # empty, shorter than k, several tiles per CTA
    table_w = synth.make_tokens_for_words(words, seed=9)
    dev = torch.device(DEV)
    store = DocStore.from_lists(words, boxes, labels, batch["page_indices"], lambda w: table_w.get(w, [2]), dev,
                                images=batch["images"], derived=derived)
    prompts = prompts_for(["what is item %d about ?" % b for b in range(len(words))])
    emb = [e.to(dev) for e in batch["text_embeddings"]]
    q = batch["question_embeddings"].to(dev)
    table = F.build_doc_table(emb, q.shape[1], dev)
    B = len(emb)

    def run(one_launch):
        sims = torch.zeros(table.total_rows, dtype=torch.float32, device=dev)
        idx = torch.full((B, k), -7, dtype=torch.int32, device=dev)
        val = torch.zeros((B, k), dtype=torch.float32, device=dev)
        cnt = torch.full((B,), -7, dtype=torch.int32, device=dev)
        plan = store.prepare_gather(idx, cnt, prompts, reorder_chunks=reorder, sep_ids=[2, 9] if sep else [], max_len=512,
                                    with_layout_labels=True, sims=sims, topk_val=val, max_rows=table.max_rows)
        outs = []
        for _ in range(2):
            if one_launch:
                assert plan.can_retrieve_in_one_launch(table)
                plan.launch_retrieve(table, q, sims)
            else:
                F.score_table(table, q, out=sims)
                plan.launch()
            pk = plan.finish()
            outs.append([t.clone() for t in (sims, idx, val, cnt, pk.input_ids, pk.boxes, pk.attention_mask, pk.layout_labels,
                                             pk.hit_chunk, pk.hit_page, pk.hit_label, pk.hit_nwords, pk.hit_bbox, pk.hit_rect)])
        for x, y in zip(*outs):
            assert torch.equal(x, y)
        return outs[0]

    for x, y in zip(run(True), run(False)):
        assert torch.equal(x, y)


def test_store_and_embeddings_must_describe_the_same_chunks():
    from rag_docvqa_b200.docstore import DocStore
    batch = synth.make_text_batch("C2", with_lists=True, docs=4, seed=3)
    table_w = synth.make_tokens_for_words(batch["words_text_chunks"], seed=9)
    dev = torch.device(DEV)
    store = DocStore.from_lists(batch["words_text_chunks"], batch["words_box_chunks"], batch["layout_labels_chunks"],
                                batch["page_indices"], lambda w: table_w.get(w, [2]), dev)
    idx = torch.empty((4, 5), dtype=torch.int32, device=dev)
    cnt = torch.empty((4,), dtype=torch.int32, device=dev)
    with pytest.raises(ValueError, match="same chunks"):
        store.prepare_gather(idx, cnt, [[1]] * 4, sims=torch.zeros(store.n_chunks + 1, device=dev),
                             topk_val=torch.empty((4, 5), device=dev))
    assert store.same_chunks_as(batch["sizes"]) and not store.same_chunks_as(batch["sizes"][::-1])


def test_resident_embedding_cache():
    """retrieval_embedding_cache_mb: a second question about the same host documents is answered from their device copies --
    same outputs as without the cache; an in-place change of a document is seen (version counter)."""
    from rag_docvqa_b200.retriever import Retriever
    batch = synth.make_text_batch("C2", with_lists=True, docs=10, seed=21)
    lists = (batch["words_text_chunks"], batch["words_box_chunks"], batch["layout_labels_chunks"], batch["images"],
             batch["page_indices"])
    emb = [e.clone().pin_memory() for e in batch["text_embeddings"]]
    q = batch["question_embeddings"]
    plain = Retriever({**BASE, "chunk_num": 5})
    cached = Retriever({**BASE, "chunk_num": 5, "retrieval_embedding_cache_mb": 64})
    want = plain.retrieve(emb, q, *lists)
    for rep in range(3):
        got = cached.retrieve(emb, q, *lists)
        for i in (0, 1, 2, 3, 4, 5, 7):
            assert got[i] == want[i]
        for a, b in zip(got[8], want[8]):
            assert torch.equal(a, b) and not a.is_cuda
    n_docs = sum(1 for e in emb if e.shape[0])
    assert cached._cache.misses == n_docs and cached._cache.hits == 2 * n_docs
    emb[0].mul_(-1.0)                                                   # in-place torch write: the version counter moves
    got = cached.retrieve(emb, q, *lists)
    want2 = plain.retrieve(emb, q, *lists)
    assert got[0] == want2[0] and torch.equal(got[8][0], want2[8][0]) and cached._cache.misses == n_docs + 1
    tiny = Retriever({**BASE, "chunk_num": 5, "retrieval_embedding_cache_mb": 0.05})       # 50 KB: eviction, still correct
    for rep in range(2):
        got = tiny.retrieve(emb, q, *lists)
        assert got[0] == want2[0]
    assert tiny._cache.bytes <= 0.05 * (1 << 20)


def test_packed_inputs_fuzz():
    """Sixteen seeded combinations nobody chose of k, neighbour window, chunk reordering, separator, truncation length and
    store construction: the packed ids / boxes / mask / labels must equal the oracle's flatten + VT5 packing bit for bit."""
    from rag_docvqa_b200.docstore import DocStore
    from rag_docvqa_b200.retriever import Retriever
    rng = np.random.RandomState(99)
    for case in range(16):
        k = int(rng.randint(1, 9))
        s = int(rng.choice([0, 0, 1, 4, 50]))
        reorder, sep, derived = bool(rng.randint(2)), bool(rng.randint(2)), bool(rng.randint(2))
        max_len = int(rng.choice([24, 65, 200, 512]))
        batch = synth.make_text_batch("C2", with_lists=True, docs=int(rng.randint(2, 7)), seed=300 + case, dup_frac=0.05)
        words, boxes, labels = batch["words_text_chunks"], batch["words_box_chunks"], batch["layout_labels_chunks"]
        pages, images = batch["page_indices"], batch["images"]
        table = synth.make_tokens_for_words(words, seed=case)
        tok = lambda w, _t=table: _t.get(w, [2])
        store = DocStore.from_lists(words, boxes, labels, pages, tok, torch.device(DEV), images=images, derived=derived)
        prompts = prompts_for(["what is item %d about ?" % b for b in range(len(words))])
        sep_ids = [2, 9] if sep else []
        retr = Retriever({**BASE, "chunk_num": k, "include_surroundings": s, "reorder_chunks": reorder})
        packed, res = retr.retrieve_packed([e.to(DEV) for e in batch["text_embeddings"]], batch["question_embeddings"].to(DEV),
                                           store, prompts, sep_ids=sep_ids, max_source_length=max_len, with_layout_labels=True)
        hits = Retriever._hits_to_host(res.topk_idx, res.topk_cnt)
        ref = R.gather_hits(hits, words, boxes, labels, images, pages, include_surroundings=s, reorder_chunks=reorder, crop=False)
        sep_word = "<sep>" if sep else None
        tok_ref = lambda w, _s=sep_ids, _t=tok: _s if w == "<sep>" else _t(w)
        ids, bxs, mask, labs = R.vt5_pack(prompts, [R.flatten(x, sep_word) for x in ref[3]], [R.flatten(x, sep_word) for x in ref[4]],
                                          tok_ref, layout_labels=[R.flatten(x, sep_word) for x in ref[5]], max_source_length=max_len)
        what = "case %d: k=%d s=%d reorder=%s sep=%s max_len=%d derived=%s" % (case, k, s, reorder, sep, max_len, derived)
        assert torch.equal(packed.input_ids.cpu(), ids), what
        assert torch.equal(packed.boxes.cpu(), bxs), what
        assert torch.equal(packed.attention_mask.cpu(), mask), what
        assert torch.equal(packed.layout_labels.cpu(), labs), what
