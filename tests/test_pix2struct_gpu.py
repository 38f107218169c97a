"""GPU parity of the Pix2Struct input assembly (rdv_pix2struct_patches) against the frozen outputs of the reference's
own image processor (tests/golden/pix2struct_patches.npz) and the oracle restatement.

fp32 path; the device kernel and torch's CPU kernel apply the same anti-aliased bilinear weights but sum in a different
order and compute the image mean in double instead of numpy's pairwise float32, so values agree to float rounding:
|delta| <= 2e-5 * max(1, |ref|) (normalised pixels are O(1)).  Row / column ids, zero padding and the mask are exact."""
import os

import numpy as np
import pytest
import torch

from oracle import ref_restated as R

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
ATOL = 2e-5


def run(images_per_doc, max_total, normalize=True, crops=None):
    from PIL import Image
    from rag_docvqa_b200.pagestore import PageStore
    pil = [[Image.fromarray(im, "RGB") for im in doc] for doc in images_per_doc]
    store = PageStore.from_images(pil, torch.device(DEV))
    if crops is None:
        crops = [[(i, 0, 0, im.shape[1], im.shape[0]) for i, im in enumerate(doc)] for doc in images_per_doc]
    out = store.pack_pix2struct(crops, max_total_patches=max_total, normalize=normalize)
    torch.cuda.synchronize()
    return out.flattened_patches.cpu().numpy(), out.attention_mask.cpu().numpy()


def check(flat, mask, ref_flat, ref_mask):
    np.testing.assert_array_equal(flat[:, :2], ref_flat[:, :2])                       # row / column ids, padding
    np.testing.assert_array_equal(mask, ref_mask)
    assert np.all(np.abs(flat - ref_flat) <= ATOL * np.maximum(1.0, np.abs(ref_flat)))


def test_matches_reference_golden(golden_dir):
    z = np.load(os.path.join(golden_dir, "pix2struct_patches.npz"))
    docs = [[z["img_%d_%d" % (b, i)] for i in range(int(z["n_%d" % b]))] for b in range(int(z["docs"]))]
    flat, mask = run(docs, 128)
    for b in range(len(docs)):
        check(flat[b], mask[b], z["flat_%d" % b], z["mask_%d" % b])


@pytest.mark.parametrize("max_total,normalize", [(2048, True), (256, True), (64, False)])
def test_matches_oracle_on_crops(max_total, normalize):
    rng = np.random.RandomState(max_total)
    pages = [[rng.randint(0, 256, (rng.randint(120, 400), rng.randint(150, 500), 3)).astype(np.uint8) for _ in range(3)] for _ in range(4)]
    crops, ref_imgs = [], []
    for b, doc in enumerate(pages):
        c_b, r_b = [], []
        for i in range(rng.randint(1, 6)):
            p = rng.randint(0, 3)
            H, W = doc[p].shape[:2]
            x0, y0 = rng.randint(0, W - 20), rng.randint(0, H - 20)
            x1, y1 = rng.randint(x0 + 8, W + 1), rng.randint(y0 + 8, H + 1)
            c_b.append((p, x0, y0, x1, y1))
            r_b.append(doc[p][y0:y1, x0:x1])
        crops.append(c_b)
        ref_imgs.append(r_b)
    flat, mask = run(pages, max_total, normalize=normalize, crops=crops)
    for b in range(len(pages)):
        ref_flat, ref_mask = R.pix2struct_patches(ref_imgs[b], max_total, normalize=normalize)
        check(flat[b], mask[b], ref_flat, ref_mask)


def test_no_images_raises_like_the_reference():
    from PIL import Image
    from rag_docvqa_b200.pagestore import PageStore
    store = PageStore.from_images([[Image.new("RGB", (40, 30))]], torch.device(DEV))
    with pytest.raises(ValueError, match="No images provided"):
        store.pack_pix2struct([[]])


def test_visual_retrieve_packed_equals_reference_processor_on_the_crops():
    """VisualRetriever.retrieve_packed = the oracle's patch assembly applied to the crops VisualRetriever.retrieve returns."""
    from PIL import Image
    from rag_docvqa_b200.pagestore import PageStore
    from rag_docvqa_b200.retriever import VisualRetriever
    rng = np.random.RandomState(4)
    dev = torch.device(DEV)
    B, strips, L, d = 3, 6, 64, 32
    images, flat, mats, xyxy = [], [], [], []
    for b in range(B):
        pages = [Image.fromarray(rng.randint(0, 256, (180, 240, 3)).astype(np.uint8), "RGB") for _ in range(2)]
        images.append(pages)
        f_b, m_b, x_b = [], [], []
        for g in range(2):                                           # 3 horizontal strips per page
            rows = [[0, 60 * r, 240, 60 * r + 60] for r in range(3)]
            x_b.append(rows)
            m_b.append([[pages[g].crop(tuple(rc))] for rc in rows])
            f_b.extend([g] * 3)
        flat.append(np.array(f_b, dtype=np.int64)); mats.append(m_b); xyxy.append(x_b)
    g_t = torch.Generator().manual_seed(3)
    patches = [torch.randn(strips, L, d, generator=g_t).to(dev) for _ in range(B)]
    q = torch.randn(B, L, d, generator=g_t).to(dev)
    vr = VisualRetriever({"chunk_num": 3, "include_surroundings": 0, "chunk_mode": "horizontal", "device": DEV})
    crops, page_ids = vr.retrieve(patches, q, flat, mats, xyxy, images)
    store = PageStore.from_images(images, dev)
    packed, page_ids2 = vr.retrieve_packed(patches, q, flat, mats, xyxy, store, max_total_patches=512)
    torch.cuda.synchronize()
    assert page_ids2 == page_ids
    for b in range(B):
        ref_flat, ref_mask = R.pix2struct_patches([np.asarray(c) for c in crops[b]], 512)
        check(packed.flattened_patches[b].cpu().numpy(), packed.attention_mask[b].cpu().numpy(), ref_flat, ref_mask)
