"""TEST INFRASTRUCTURE ONLY -- freezes the UNMODIFIED reference's Pix2Struct input assembly
(src/custom_pix2struct_processor.py: CustomPix2StructImageProcessor.normalize + extract_multi_image_flattened_patches +
the attention mask of preprocess) as tests/golden/pix2struct_patches.npz.

    PYTHONDONTWRITEBYTECODE=1 python -m oracle.make_golden_pix2struct     (build container: /root/reference must exist)

Two environment shims, neither touches the reference's code: transformers 5.x dropped `render_header` (imported at
:10-13, used only for the header text, out of scope) and gave `torch_extract_patches` a batch dimension; the
reference is pinned to transformers==4.49.0, whose function takes (C, H, W) -- the shim adds / strips that dimension.
"""
import importlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.ref_import import REFERENCE_ROOT  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def import_reference_processor():
    import transformers.models.pix2struct.image_processing_pix2struct as ip
    sys.dont_write_bytecode = True
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    if not hasattr(ip, "render_header"):
        def _no_header(*a, **k):
            raise RuntimeError("render_header is out of scope")
        ip.render_header = _no_header
    sys.modules.pop("src.custom_pix2struct_processor", None)
    mod = importlib.import_module("src.custom_pix2struct_processor")
    new_fn = ip.torch_extract_patches
    import torch
    try:
        new_fn(torch.zeros(3, 16, 16), 16, 16)                        # transformers 4.49 signature: (C, H, W)
    except ValueError:                                                # transformers 5.x wants (B, C, H, W)
        mod.torch_extract_patches = lambda img, ph, pw: new_fn(img.unsqueeze(0), ph, pw)
    return mod


def make_inputs(seed=77):
    rng = np.random.RandomState(seed)
    docs = []
    for n in (1, 3, 5):
        docs.append([rng.randint(0, 256, (rng.randint(20, 90), rng.randint(40, 200), 3)).astype(np.uint8) for _ in range(n)])
    docs.append([np.full((33, 47, 3), 200, np.uint8)])             # constant image: std floor 1/sqrt(#elements)
    return docs


def main():
    mod = import_reference_processor()
    proc = mod.CustomPix2StructImageProcessor(max_total_patches=128, is_vqa=False)
    out = {"docs": np.int64(0)}
    docs = make_inputs()
    for b, images in enumerate(docs):
        normed = [proc.normalize(image=im, input_data_format="channels_last") for im in images]
        flat = proc.extract_flattened_patches(normed, input_data_format="channels_last")
        out["flat_%d" % b] = flat.astype(np.float32)
        out["mask_%d" % b] = (flat.sum(axis=-1) != 0).astype(np.float32)
        out["n_%d" % b] = np.int64(len(images))
        for i, im in enumerate(images):
            out["img_%d_%d" % (b, i)] = im
    out["docs"] = np.int64(len(docs))
    np.savez_compressed(os.path.join(GOLDEN, "pix2struct_patches.npz"), **out)
    mpath = os.path.join(GOLDEN, "MANIFEST.json")
    manifest = json.load(open(mpath))
    import torch, transformers
    manifest["files"]["pix2struct_patches.npz"] = (
        "CustomPix2StructImageProcessor.normalize + extract_multi_image_flattened_patches (src/custom_pix2struct_processor.py:"
        "97-132, 175-196), max_total_patches=128; torch %s, transformers %s (torch_extract_patches shimmed to the 4.49 signature)"
        % (torch.__version__, transformers.__version__))
    json.dump(manifest, open(mpath, "w"), indent=1)
    print("wrote pix2struct_patches.npz", os.path.getsize(os.path.join(GOLDEN, "pix2struct_patches.npz")), "bytes")


if __name__ == "__main__":
    main()
