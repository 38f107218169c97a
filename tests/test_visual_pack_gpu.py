"""GPU parity of the device path from crop rectangles to the generator's visual input (rdv_visual_pack) against the
oracle's restatement of concatenate_patches(mode="grid") + Pillow's resize, which tests/test_oracle_golden.py pins to
the reference / the installed Pillow.  uint8 results must be bit-exact."""
import numpy as np
import pytest
import torch

from oracle import ref_restated as R

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def make_case(seed, B, k, max_wh=(300, 260), full_page_every=0, out_of_page=False):
    from PIL import Image
    rng = np.random.RandomState(seed)
    images, pages_np = [], []
    for b in range(B):
        n_pages = rng.randint(1, 4)
        arrs = [rng.randint(0, 256, (rng.randint(40, max_wh[1]), rng.randint(40, max_wh[0]), 3)).astype(np.uint8)
                for _ in range(n_pages)]
        pages_np.append(arrs)
        images.append([Image.fromarray(a, "RGB") for a in arrs])
    hit_page = np.full((B, k), -1, np.int32)
    hit_rect = np.full((B, k, 4), -1, np.int32)
    hit_cnt = np.zeros((B,), np.int32)
    for b in range(B):
        n = rng.randint(0, k + 1) if b else k
        hit_cnt[b] = n
        for i in range(n):
            p = rng.randint(0, len(pages_np[b]))
            H, W = pages_np[b][p].shape[:2]
            if full_page_every and i % full_page_every == 0:
                rect = (0, 0, W, H)
            else:
                x0, y0 = rng.randint(0, W - 1), rng.randint(0, H - 1)
                x1, y1 = rng.randint(x0 + 1, W + 1), rng.randint(y0 + 1, H + 1)
                if out_of_page and i % 2:
                    x0, y1 = x0 - 7, y1 + 9            # PIL crops are black outside the page
                rect = (x0, y0, x1, y1)
            hit_page[b, i] = p
            hit_rect[b, i] = rect
    return images, pages_np, hit_page, hit_rect, hit_cnt


def run(images, hit_page, hit_rect, hit_cnt, **kw):
    from rag_docvqa_b200.pagestore import PageStore
    store = PageStore.from_images(images, torch.device(DEV))
    out = store.pack(torch.from_numpy(hit_page).to(DEV), torch.from_numpy(hit_rect).to(DEV),
                     torch.from_numpy(hit_cnt).to(DEV), **kw)
    torch.cuda.synchronize()
    return out


@pytest.mark.parametrize("resample", [R.PIL_BILINEAR, R.PIL_BICUBIC])
@pytest.mark.parametrize("seed,B,k,out_size,kw", [
    (1, 6, 5, 224, {}), (2, 4, 8, 224, dict(full_page_every=2)), (3, 5, 3, 64, dict(out_of_page=True)),
    (4, 3, 20, 224, dict(max_wh=(120, 90))), (5, 2, 1, 224, dict(max_wh=(60, 50))),        # up-scaling
    (6, 3, 4, 50, {})])                     # 50 * 3 bytes per row is not a multiple of 4: the strided vertical pass
def test_visual_pack_matches_oracle(seed, B, k, out_size, kw, resample):
    images, pages_np, hit_page, hit_rect, hit_cnt = make_case(seed, B, k, **kw)
    mean, std = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)
    out = run(images, hit_page, hit_rect, hit_cnt, out_size=out_size, resample=resample, mean=mean, std=std)
    got_u8, got_px, status = out.image_u8.cpu().numpy(), out.pixel_values.cpu().numpy(), out.status.cpu().numpy()
    for b in range(B):
        n = hit_cnt[b]
        rects = [tuple(int(v) for v in hit_rect[b, i]) for i in range(n)]
        sizes = [(r[2] - r[0], r[3] - r[1]) for r in rects]
        gw, gh, _ = R.grid_layout(sizes)
        if gw <= 0 or gh <= 0:
            assert status[b] == 2
            continue
        assert status[b] == 0
        ref_u8, ref_px = R.visual_input(pages_np[b], rects, [int(p) for p in hit_page[b, :n]], out_size, resample, mean, std)
        np.testing.assert_array_equal(got_u8[b], ref_u8, err_msg="doc %d" % b)
        np.testing.assert_allclose(got_px[b], ref_px, rtol=1e-6, atol=1e-6)


def test_visual_pack_matches_pillow_directly():
    """End to end against the real thing: PIL crop -> paste canvas -> PIL resize."""
    from PIL import Image
    images, pages_np, hit_page, hit_rect, hit_cnt = make_case(11, 4, 5)
    out = run(images, hit_page, hit_rect, hit_cnt, resample=R.PIL_BICUBIC)
    got = out.image_u8.cpu().numpy()
    for b in range(4):
        n = hit_cnt[b]
        patches = [images[b][hit_page[b, i]].crop(tuple(int(v) for v in hit_rect[b, i])) for i in range(n)]
        sizes = [p.size for p in patches]
        gw, gh, pos = R.grid_layout(sizes)
        canvas = Image.new("RGB", (gw, gh))
        for p, xy in zip(patches, pos):
            canvas.paste(p, xy)
        ref = np.asarray(canvas.resize((224, 224), resample=Image.Resampling.BICUBIC))
        np.testing.assert_array_equal(got[b], ref)


def test_visual_pack_after_gather_c2_slice():
    """The whole packed path: score -> select + gather (hit_page / hit_rect on the device) -> visual pack."""
    from rag_docvqa_b200 import synth
    from rag_docvqa_b200.docstore import DocStore
    from rag_docvqa_b200.pagestore import PageStore
    from rag_docvqa_b200.retriever import Retriever
    batch = synth.make_text_batch("C2", with_lists=True, docs=6, seed=5, dup_frac=0.0)
    words, boxes, labels = batch["words_text_chunks"], batch["words_box_chunks"], batch["layout_labels_chunks"]
    pages, images = batch["page_indices"], batch["images"]
    table = synth.make_tokens_for_words(words, seed=9)
    dev = torch.device(DEV)
    store = DocStore.from_lists(words, boxes, labels, pages, lambda w: table.get(w, [2]), dev, images=images)
    pstore = PageStore.from_images(images, dev)
    retr = Retriever({"compute_stats": False, "compute_stats_examples": False, "n_stats_examples": 0, "device": DEV, "chunk_num": 5})
    prompts = [[5, 6, 7]] * len(words)
    packed, res = retr.retrieve_packed([e.to(dev) for e in batch["text_embeddings"]], batch["question_embeddings"].to(dev),
                                       store, prompts)
    vis = pstore.pack(packed.hit_page, packed.hit_rect, res.topk_cnt)
    torch.cuda.synchronize()
    hits = Retriever._hits_to_host(res.topk_idx, res.topk_cnt)
    ref = R.gather_hits(hits, words, boxes, labels, images, pages, include_surroundings=0, reorder_chunks=False, crop=False)
    for b in range(len(words)):
        if not hits[b]:
            continue
        pages_np = [np.asarray(im) for im in images[b]]
        ref_u8, _ = R.visual_input(pages_np, [tuple(r) for r in ref[6][b]], ref[7][b])
        np.testing.assert_array_equal(vis.image_u8[b].cpu().numpy(), ref_u8)


@pytest.mark.parametrize("name,resample", [("bilinear", R.PIL_BILINEAR), ("bicubic", R.PIL_BICUBIC)])
def test_visual_pack_matches_reference_golden(golden_dir, name, resample):
    """tests/golden/visual_pack.npz: the reference's own concatenate_patches(mode="grid") + PIL resize, frozen."""
    import os
    from PIL import Image
    z = np.load(os.path.join(golden_dir, "visual_pack.npz"))
    docs, k = int(z["docs"]), 6
    images, hit_page = [], np.full((docs, k), -1, np.int32)
    hit_rect, hit_cnt = np.full((docs, k, 4), -1, np.int32), np.zeros((docs,), np.int32)
    for b in range(docs):
        images.append([Image.fromarray(z["page_%d_%d" % (b, p)], "RGB") for p in range(int(z["n_pages_%d" % b]))])
        n = len(z["page_of_%d" % b])
        hit_cnt[b] = n
        hit_page[b, :n] = z["page_of_%d" % b]
        hit_rect[b, :n] = z["rects_%d" % b]
    out = run(images, hit_page, hit_rect, hit_cnt, out_size=64, resample=resample)
    assert out.status.cpu().tolist() == [0] * docs
    for b in range(docs):
        np.testing.assert_array_equal(out.image_u8[b].cpu().numpy(), z["resized_%s_%d" % (name, b)], err_msg="doc %d" % b)
