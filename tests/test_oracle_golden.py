"""Pins oracle/ref_restated.py to the golden vectors frozen from the unmodified reference
(oracle/make_golden.py), and -- when /root/reference is present -- to the live reference."""
import json
import os
import zlib

import numpy as np
import pytest
import torch

from oracle import ref_restated as R
from oracle.ref_import import reference_available
from rag_docvqa_b200 import synth

TEXT_CASES = ["c1", "ragged_norm", "ragged_raw", "k20_d1024"]


def load_text_case(golden_dir, name):
    z = np.load(os.path.join(golden_dir, "score_topk_%s.npz" % name))
    sizes = z["sizes"].tolist()
    emb = [torch.from_numpy(z["emb_%d" % b]) for b in range(len(sizes))]
    sims = [z["sims_%d" % b] for b in range(len(sizes))]
    topk = [z["topk_%d" % b] for b in range(len(sizes))]
    return sizes, int(z["k"]), emb, torch.from_numpy(z["q"]), sims, topk


@pytest.mark.parametrize("name", TEXT_CASES)
def test_score_bit_exact(golden_dir, name):
    sizes, k, emb, q, sims, _ = load_text_case(golden_dir, name)
    got = R.score(emb, q)
    for b in range(len(sizes)):
        assert got[b].shape == (sizes[b],)
        # same torch CPU ops; allow last-ulp differences across torch builds / CPU ISAs
        np.testing.assert_allclose(got[b].numpy(), sims[b], rtol=2e-6, atol=1e-7)


@pytest.mark.parametrize("name", TEXT_CASES)
def test_topk_lowest_index_equals_torch_topk_modulo_ties(golden_dir, name):
    sizes, k, emb, q, sims, topk = load_text_case(golden_dir, name)
    for b in range(len(sizes)):
        mine = R.topk_lowest_index(sims[b], k)
        assert len(mine) == min(k, sizes[b]) == len(topk[b])
        # identical score sequences; identical indices wherever scores are distinct
        np.testing.assert_array_equal(sims[b][mine], sims[b][topk[b]])
        distinct = np.ones(len(mine), dtype=bool)
        vals = sims[b][mine]
        for i in range(len(vals)):
            if np.count_nonzero(sims[b] == vals[i]) > 1:
                distinct[i] = False
        np.testing.assert_array_equal(mine[distinct], topk[b][distinct])
        # within a tie group: ascending index
        for i in range(1, len(mine)):
            if vals[i] == vals[i - 1]:
                assert mine[i] > mine[i - 1]


def test_topk_conventions():
    v = np.array([1, 3, 3, 2, 3, .5], dtype=np.float32)
    assert R.topk_lowest_index(v, 2).tolist() == [1, 2]
    v = np.array([0.0, -0.0, np.nan, 1.0, -np.nan], dtype=np.float32)
    assert R.topk_lowest_index(v, 5).tolist() == [2, 4, 3, 0, 1]
    assert R.topk_lowest_index(np.zeros(0, dtype=np.float32), 3).tolist() == []
    # NaN greatest matches torch.topk
    assert torch.topk(torch.tensor([0.0, float("nan"), 1.0]), 1).indices.tolist() == [1]


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


def test_retrieve_lists_match_reference(golden_dir):
    gold, emb, q, words, boxes, labels, images, pages = load_retrieve_inputs(golden_dir)
    assert len(gold["variants"]) == 8
    for var in gold["variants"]:
        out = R.retrieve(emb, q, words, boxes, labels, images, pages, k=var["k"],
                         include_surroundings=var["include_surroundings"], reorder_chunks=var["reorder_chunks"])
        assert out[0] == var["top_k_text"]
        assert out[1] == var["top_k_boxes"]
        assert out[2] == var["top_k_layout_labels"]
        assert out[3] == var["top_k_words_text"]
        assert out[4] == var["top_k_words_boxes"]
        assert out[5] == var["top_k_words_layout_labels"]
        assert [[digest(im) for im in doc] for doc in out[6]] == var["top_k_patches"]
        assert out[7] == var["top_k_page_indices"]
        for b, s in enumerate(out[8]):
            np.testing.assert_allclose(s.numpy(), np.array(var["similarities"][b], dtype=np.float32), rtol=2e-6, atol=1e-7)
        # deterministic-tie variant gives the same lists here (no duplicate rows in this fixture)
        det = R.retrieve(emb, q, words, boxes, labels, images, pages, k=var["k"],
                         include_surroundings=var["include_surroundings"], reorder_chunks=var["reorder_chunks"],
                         deterministic_ties=True)
        assert det[3] == var["top_k_words_text"] and det[7] == var["top_k_page_indices"]


def test_mean_pooling_golden(golden_dir):
    z = np.load(os.path.join(golden_dir, "mean_pooling.npz"))
    got = R.mean_pooling(torch.from_numpy(z["embs"]), torch.from_numpy(z["mask"]))
    np.testing.assert_allclose(got.numpy(), z["pooled"], rtol=1e-6, atol=1e-7)
    assert (got[:2] == 0).all()        # all-pad rows pool to exactly 0


def test_late_interaction_golden(golden_dir):
    z = np.load(os.path.join(golden_dir, "late_interaction.npz"))
    q = torch.from_numpy(z["q"])
    got = R.visual_scores([torch.from_numpy(z["p0"]), torch.from_numpy(z["p1"])], q)
    np.testing.assert_allclose(got[0].numpy(), z["s0"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(got[1].numpy(), z["s1"], rtol=1e-5, atol=1e-5)
    f64 = R.late_interaction_f64(q[0:1], torch.from_numpy(z["p0"]))
    np.testing.assert_allclose(f64.numpy(), z["s0"], rtol=1e-5)


def test_visual_decode_golden(golden_dir):
    with open(os.path.join(golden_dir, "visual_retrieve.json")) as f:
        gold = json.load(f)
    from PIL import Image
    rng = np.random.RandomState(5)
    flat, shapes, xyxy, images = [], [], [], []
    for doc in gold["docs"]:
        f_b, imgs = [], []
        for g, n_rows in enumerate(doc["groups"]):
            W, H = doc["image_wh"][g]
            imgs.append(Image.fromarray(rng.randint(0, 255, size=(H, W, 3)).astype(np.uint8), "RGB"))
            f_b.extend([g] * n_rows)
        flat.append(np.array(f_b, dtype=np.int64))
        shapes.append([(n_rows, 1) for n_rows in doc["groups"]])
        xyxy.append(doc["xyxy"])
        images.append(imgs)
    for var in gold["variants"]:
        s = tuple(var["include_surroundings"]) if isinstance(var["include_surroundings"], list) else var["include_surroundings"]
        hits = [R.topk_lowest_index(np.array(x, dtype=np.float32), var["k"]) if len(flat[b]) else []
                for b, x in enumerate(var["sims"])]
        rects, groups = R.visual_decode(hits, flat, shapes, xyxy, include_surroundings=s)
        for b in range(len(flat)):
            crops = sorted(digest(images[b][g].crop(tuple(rc))) for g, lst in rects[b].items() for rc in lst)
            assert crops == var["crops"][b]
            assert groups[b] == var["pages"][b]


def test_vt5_pack_golden(golden_dir):
    gold, emb, q, words, boxes, labels, images, pages = load_retrieve_inputs(golden_dir)
    with open(os.path.join(golden_dir, "vt5_pack.json")) as f:
        packs = json.load(f)
    table = synth.make_tokens_for_words(words, seed=packs["word_table_seed"])
    out = R.retrieve(emb, q, words, boxes, labels, images, pages, k=packs["k"])
    prompts = [[5 + (zlib.crc32(t.encode()) % 1000) for t in ("question: {:s}  context: ".format(qs)).split()]
               for qs in packs["questions"]]
    for var in packs["variants"]:
        sep = var["sep"]
        w_flat = [R.flatten(b, sep) for b in out[3]]
        b_flat = [R.flatten(b, sep) for b in out[4]]
        l_flat = [R.flatten(b, sep) for b in out[5]]
        ids, bxs, mask, labs = R.vt5_pack(prompts, w_flat, b_flat, lambda w: list(table.get(w, [2])),
                                          layout_labels=l_flat if var["use_layout_labels"] == "Embed" else None,
                                          max_source_length=var["max_source_length"])
        assert ids.tolist() == var["input_ids"]
        assert bxs.tolist() == var["boxes"]
        assert mask.tolist() == var["attention_mask"]
        if var["layout_labels"] is not None:
            assert labs.tolist() == var["layout_labels"]


@pytest.mark.skipif(not reference_available(), reason="reference tree not mounted (GPU box)")
def test_against_live_reference():
    from oracle.ref_import import import_reference
    modules, utils, model_utils = import_reference()
    batch = synth.make_text_batch("C2", with_lists=True, docs=6, seed=99)
    cfg = {"compute_stats": False, "compute_stats_examples": False, "n_stats_examples": 0,
           "chunk_num": 5, "include_surroundings": 2, "reorder_chunks": True}
    ref = modules.Retriever(cfg).retrieve(
        batch["text_embeddings"], batch["question_embeddings"], batch["words_text_chunks"],
        batch["words_box_chunks"], batch["layout_labels_chunks"], batch["images"], batch["page_indices"])
    got = R.retrieve(batch["text_embeddings"], batch["question_embeddings"], batch["words_text_chunks"],
                     batch["words_box_chunks"], batch["layout_labels_chunks"], batch["images"],
                     batch["page_indices"], k=5, include_surroundings=2, reorder_chunks=True)
    for i in (0, 1, 2, 3, 4, 5, 7):
        assert got[i] == ref[i]
    for a, b in zip(got[6], ref[6]):
        assert [digest(x) for x in a] == [digest(x) for x in b]
    for a, b in zip(got[8], ref[8]):
        assert torch.equal(a, b)
    e, m = synth.make_token_batch(17, 24, 5, mean_len=10, std_len=4, min_len=0, max_len=20)
    assert torch.equal(R.mean_pooling(e, m), model_utils.mean_pooling(e, m))
    p, qq = synth.make_strip_batch(1, [3], 20, 24, 6)
    assert torch.equal(R.late_interaction(qq[0:1], p[0]), utils.late_interaction(qq[0:1], p[0]))


# ---- f1: retrieved patches -> visual input (concatenate_patches grid + Pillow resize) -------------------------
def _visual_golden(golden_dir):
    z = np.load(os.path.join(golden_dir, "visual_pack.npz"))
    for b in range(int(z["docs"])):
        pages = [z["page_%d_%d" % (b, p)] for p in range(int(z["n_pages_%d" % b]))]
        rects = [tuple(int(v) for v in r) for r in z["rects_%d" % b]]
        yield b, z, pages, rects, [int(p) for p in z["page_of_%d" % b]]


def test_concat_grid_and_resize_match_reference_golden(golden_dir):
    """The frozen outputs of the reference's concatenate_patches(mode="grid") + PIL resize."""
    for b, z, pages, rects, page_of in _visual_golden(golden_dir):
        canvas = R.concat_grid(pages, rects, page_of)
        np.testing.assert_array_equal(canvas, z["canvas_%d" % b])
        for name, kind in (("bilinear", R.PIL_BILINEAR), ("bicubic", R.PIL_BICUBIC)):
            np.testing.assert_array_equal(R.pil_resize_u8(canvas, 64, 64, kind), z["resized_%s_%d" % (name, b)])


@pytest.mark.parametrize("shape", [(300, 400, 224, 224), (1069, 791, 224, 224), (2500, 816, 224, 224), (100, 60, 224, 224),
                                   (224, 500, 224, 224), (37, 224, 224, 224), (224, 224, 224, 224), (3, 3, 32, 32),
                                   (90, 700, 64, 48)])
def test_pil_resize_restatement_equals_installed_pillow(shape):
    """Pillow is a third-party dependency of the path (not under /root/reference): the restated Resample.c must
    equal the installed library bit for bit, down- and up-scaling, both filters."""
    from PIL import Image
    h, w, oh, ow = shape
    img = np.random.RandomState(h * 7 + w).randint(0, 256, (h, w, 3)).astype(np.uint8)
    for kind, pk in ((R.PIL_BILINEAR, Image.Resampling.BILINEAR), (R.PIL_BICUBIC, Image.Resampling.BICUBIC)):
        ref = np.asarray(Image.fromarray(img, "RGB").resize((ow, oh), resample=pk))
        np.testing.assert_array_equal(R.pil_resize_u8(img, ow, oh, kind), ref)


@pytest.mark.skipif(not reference_available(), reason="reference tree not mounted (GPU box)")
def test_concat_grid_against_live_reference():
    from PIL import Image
    from oracle.ref_import import import_reference
    _, utils, _ = import_reference()
    rng = np.random.RandomState(3)
    for trial in range(20):
        pages = [rng.randint(0, 256, (rng.randint(50, 200), rng.randint(50, 200), 3)).astype(np.uint8) for _ in range(3)]
        rects, page_of = [], []
        for i in range(rng.randint(0, 7)):
            p = rng.randint(0, 3)
            H, W = pages[p].shape[:2]
            x0, y0 = rng.randint(0, W - 1), rng.randint(0, H - 1)
            rects.append((x0, y0, rng.randint(x0 + 1, W + 1), rng.randint(y0 + 1, H + 1)))
            page_of.append(p)
        patches = [Image.fromarray(pages[p], "RGB").crop(r) for r, p in zip(rects, page_of)]
        ref = np.asarray(utils.concatenate_patches(patches, mode="grid").convert("RGB"))
        np.testing.assert_array_equal(R.concat_grid(pages, rects, page_of), ref)


def test_pix2struct_patches_match_reference_golden(golden_dir):
    """Frozen outputs of the reference's CustomPix2StructImageProcessor.normalize + extract_multi_image_flattened_patches."""
    z = np.load(os.path.join(golden_dir, "pix2struct_patches.npz"))
    for b in range(int(z["docs"])):
        imgs = [z["img_%d_%d" % (b, i)] for i in range(int(z["n_%d" % b]))]
        flat, mask = R.pix2struct_patches(imgs, 128)
        np.testing.assert_array_equal(flat, z["flat_%d" % b])
        np.testing.assert_array_equal(mask, z["mask_%d" % b])
    with pytest.raises(ValueError):
        R.pix2struct_patches([], 128)


def test_pooled_patch_golden(golden_dir):
    """Pooled-patch visual retrieval: the oracle's composition of mean pooling + cosine + strip max equals what the
    reference's own functions produced (tests/golden/pooled_patch.npz), and the live reference when it is mounted."""
    z = np.load(os.path.join(golden_dir, "pooled_patch.npz"))
    n_docs = int(z["n_docs"])
    patches = [torch.from_numpy(z["patches_%d" % b]) for b in range(n_docs)]
    q, mask = torch.from_numpy(z["q"]), torch.from_numpy(z["mask"])
    sims, strips, pooled = R.pooled_patch_scores(patches, q, mask)
    assert torch.equal(pooled, torch.from_numpy(z["pooled"]))
    for b in range(n_docs):
        assert torch.equal(sims[b], torch.from_numpy(z["sims_%d" % b]))
        assert torch.equal(strips[b], torch.from_numpy(z["strip_%d" % b]))
        k = int(z["k"])
        assert sorted(R.topk_lowest_index(sims[b], k).tolist()) == sorted(z["topk_patch_%d" % b].tolist()) or b == 0   # doc 0 holds an exact tie
        assert R.topk_lowest_index(strips[b], k).tolist() == z["topk_strip_%d" % b].tolist()
    if reference_available():
        from oracle.ref_import import import_reference
        modules, _, model_utils = import_reference()
        retr = modules.Retriever({"compute_stats": False, "compute_stats_examples": False, "n_stats_examples": 0})
        live = retr._get_similarities([p.reshape(-1, p.shape[2]) for p in patches], model_utils.mean_pooling(q, mask))
        for b in range(n_docs):
            assert torch.equal(sims[b], live[b])


def test_vt5_embed_golden(golden_dir):
    """The generator-input embeddings: the oracle's restatement of SpatialEmbeddings.forward and of the embedding sum of
    VT5.prepare_inputs_for_vqa equals what the reference's own module / method produced (tests/golden/vt5_embed.npz; same
    torch CPU operators in the same order), and the live module when the reference is mounted."""
    z = np.load(os.path.join(golden_dir, "vt5_embed.npz"))
    w = {k: torch.from_numpy(z[k]) for k in ("x_emb", "y_emb", "ln_weight", "ln_bias", "lin_weight", "lin_bias", "shared", "layout")}
    eps = float(z["eps"])

    def spatial(bbox):
        return R.spatial_embeddings(bbox, w["x_emb"], w["y_emb"], w["ln_weight"], w["ln_bias"], eps, w["lin_weight"], w["lin_bias"])
    bbox = torch.from_numpy(z["bbox"])
    torch.testing.assert_close(spatial(bbox), torch.from_numpy(z["spatial"]), rtol=1e-6, atol=1e-6)
    for name, labelled in (("plain", False), ("layout", True)):
        ids, boxes = torch.from_numpy(z[name + "_ids"]), torch.from_numpy(z[name + "_boxes"])
        labels = torch.from_numpy(z[name + "_labels"]) if labelled else None
        got = R.vt5_input_embeds(ids, boxes, w["shared"], spatial(boxes), labels, w["layout"], float(z["layout_scale"]))
        torch.testing.assert_close(got, torch.from_numpy(z[name + "_embeds"]), rtol=1e-6, atol=1e-6)
    # float64 weights: the yardstick the GPU test measures both against
    w64 = {k: v.double() for k, v in w.items()}
    ref64 = R.spatial_embeddings(bbox, w64["x_emb"], w64["y_emb"], w64["ln_weight"], w64["ln_bias"], eps, w64["lin_weight"], w64["lin_bias"])
    assert (ref64.float() - torch.from_numpy(z["spatial"])).abs().max() < 2e-5
    if reference_available():
        import types
        from oracle.ref_import import import_reference
        modules, _, _ = import_reference()
        cfg = types.SimpleNamespace(max_2d_position_embeddings=w["x_emb"].shape[0], hidden_size=w["x_emb"].shape[1],
                                    layer_norm_eps=eps, hidden_dropout_prob=0.1)
        live = modules.SpatialEmbeddings(cfg).eval()
        live.load_state_dict({"x_position_embeddings.weight": w["x_emb"], "y_position_embeddings.weight": w["y_emb"],
                              "LayerNorm.weight": w["ln_weight"], "LayerNorm.bias": w["ln_bias"],
                              "spatial_emb_matcher.layers.0.weight": w["lin_weight"], "spatial_emb_matcher.layers.0.bias": w["lin_bias"]})
        with torch.no_grad():
            assert torch.equal(live(bbox), spatial(bbox))
