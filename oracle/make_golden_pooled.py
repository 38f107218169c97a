"""TEST INFRASTRUCTURE ONLY -- freezes pooled-patch visual retrieval as the UNMODIFIED reference's own functions compute it:
mean_pooling (src/_model_utils.py:49-61) of the question tokens, Retriever._get_similarities (src/_modules.py:1978-1997) of
every patch vector against the pooled question, torch.max per strip, torch.topk (src/_modules.py:2408) over patches and over
strips -> tests/golden/pooled_patch.npz.

    PYTHONDONTWRITEBYTECODE=1 python -m oracle.make_golden_pooled        (build container: /root/reference must exist)
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.ref_import import import_reference  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def make_inputs(seed=77, strips=(6, 1, 0, 3), L=48, Lq=20, d=96):
    g = torch.Generator().manual_seed(seed)
    u = torch.randn(d, generator=g)
    patches = [torch.randn(n, L, d, generator=g) + 0.5 * u for n in strips]
    patches[0][2, 7] = patches[0][4, 30]                    # an exact duplicate patch across strips: lowest index first
    patches[3][1, 5] = 0.0                                  # a zero patch scores exactly 0
    q = torch.randn(len(strips), Lq, d, generator=g) + 0.5 * u
    mask = torch.ones(len(strips), Lq, dtype=torch.int64)
    mask[1, 12:] = 0                                        # a padded question
    mask[3, 1:] = 0
    return patches, q, mask


def main():
    modules, _, model_utils = import_reference()
    patches, q, mask = make_inputs()
    retr = modules.Retriever({"compute_stats": False, "compute_stats_examples": False, "n_stats_examples": 0, "chunk_num": 5})
    pooled = model_utils.mean_pooling(q, mask)                                  # the reference, unmodified
    d = q.shape[2]
    sims = retr._get_similarities([p.reshape(-1, d) for p in patches], pooled)  # the reference, unmodified
    out = {"n_docs": np.int64(len(patches)), "q": q.numpy(), "mask": mask.numpy(), "pooled": pooled.numpy(), "k": np.int64(5)}
    for b, p in enumerate(patches):
        out["patches_%d" % b] = p.numpy()
        out["sims_%d" % b] = sims[b].numpy()
        strip = sims[b].reshape(p.shape[0], -1).max(dim=1).values if p.shape[0] else torch.empty(0)
        out["strip_%d" % b] = strip.numpy()
        out["topk_patch_%d" % b] = torch.topk(sims[b], min(5, sims[b].shape[0])).indices.numpy()
        out["topk_strip_%d" % b] = torch.topk(strip, min(5, strip.shape[0])).indices.numpy()
    np.savez_compressed(os.path.join(GOLDEN, "pooled_patch.npz"), **out)
    mpath = os.path.join(GOLDEN, "MANIFEST.json")
    manifest = json.load(open(mpath))
    manifest["files"]["pooled_patch.npz"] = ("mean_pooling (src/_model_utils.py:49-61) + Retriever._get_similarities "
                                             "(src/_modules.py:1978-1997) of the reference on every patch vector, torch.max per "
                                             "strip, torch.topk; torch %s" % torch.__version__)
    json.dump(manifest, open(mpath, "w"), indent=1)
    print("wrote pooled_patch.npz", os.path.getsize(os.path.join(GOLDEN, "pooled_patch.npz")), "bytes")


if __name__ == "__main__":
    main()
