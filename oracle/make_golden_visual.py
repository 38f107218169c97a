"""TEST INFRASTRUCTURE ONLY -- freezes the UNMODIFIED reference's concatenate_patches(mode="grid") (src/utils.py:180-231)
followed by Pillow's resize (what the HF feature extractor calls at src/_modules.py:133) as tests/golden/visual_pack.npz.

    PYTHONDONTWRITEBYTECODE=1 python -m oracle.make_golden_visual        (build container: /root/reference must exist)
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.ref_import import import_reference  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def make_inputs(seed=2024, docs=5, k=6):
    rng = np.random.RandomState(seed)
    pages, rects, page_of = [], [], []
    for b in range(docs):
        arrs = [rng.randint(0, 256, (rng.randint(30, 140), rng.randint(30, 160), 3)).astype(np.uint8) for _ in range(rng.randint(1, 4))]
        n = 0 if b == 3 else rng.randint(1, k + 1)               # document 3: no hit -> the 5 x 5 blank image
        r_b, p_b = [], []
        for i in range(n):
            p = rng.randint(0, len(arrs))
            H, W = arrs[p].shape[:2]
            x0, y0 = rng.randint(0, W - 1), rng.randint(0, H - 1)
            x1, y1 = rng.randint(x0 + 1, W + 1), rng.randint(y0 + 1, H + 1)
            if b == 1 and i == 0:
                x0, y0, x1, y1 = 0, 0, W, H                        # a whole-page crop (empty chunk, src/_modules.py:1126)
            if b == 2 and i == 1:
                x0, y1 = x0 - 5, y1 + 4                            # reaches outside the page: PIL pads with black
            r_b.append((x0, y0, x1, y1))
            p_b.append(p)
        pages.append(arrs); rects.append(r_b); page_of.append(p_b)
    return pages, rects, page_of


def main():
    from PIL import Image
    _, utils, _ = import_reference()
    pages, rects, page_of = make_inputs()
    out = {}
    for b in range(len(pages)):
        pil_pages = [Image.fromarray(a, "RGB") for a in pages[b]]
        patches = [pil_pages[p].crop(r) for r, p in zip(rects[b], page_of[b])]
        canvas = utils.concatenate_patches(patches, mode="grid")          # the reference, unmodified
        out["canvas_%d" % b] = np.asarray(canvas.convert("RGB"))
        for name, res in (("bilinear", Image.Resampling.BILINEAR), ("bicubic", Image.Resampling.BICUBIC)):
            out["resized_%s_%d" % (name, b)] = np.asarray(canvas.resize((64, 64), resample=res))
        out["n_pages_%d" % b] = np.int64(len(pages[b]))
        for p, a in enumerate(pages[b]):
            out["page_%d_%d" % (b, p)] = a
        out["rects_%d" % b] = np.asarray(rects[b], dtype=np.int32).reshape(-1, 4)
        out["page_of_%d" % b] = np.asarray(page_of[b], dtype=np.int32)
    out["docs"] = np.int64(len(pages))
    np.savez_compressed(os.path.join(GOLDEN, "visual_pack.npz"), **out)
    mpath = os.path.join(GOLDEN, "MANIFEST.json")
    manifest = json.load(open(mpath))
    import PIL
    manifest["files"]["visual_pack.npz"] = ("concatenate_patches(mode='grid') of the reference (src/utils.py:180-231) + "
                                            "PIL.Image.resize to 64 x 64, Pillow %s" % PIL.__version__)
    json.dump(manifest, open(mpath, "w"), indent=1)
    print("wrote visual_pack.npz", os.path.getsize(os.path.join(GOLDEN, "visual_pack.npz")), "bytes")


if __name__ == "__main__":
    main()
