"""TEST INFRASTRUCTURE ONLY -- freezes the reference's S2Chunker (src/_modules.py:1669-1962).

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden_s2chunker.py

Runs the UNMODIFIED reference S2Chunker (needs /root/reference) on seeded synthetic pages (rag_docvqa_b200.synth.
make_s2_pages: the tests regenerate the inputs from the seed) and writes tests/golden/s2chunker.json:
  * per page, the reference's create_nodes_and_edges node ids / used mask and _combined_weights matrix (float64, stored
    as hex so the comparison is bit-exact) in cluster_mode "spatial" and "spatial+semantic" (embedder = synth.HashEmbedder);
  * the cluster arrays S2Chunker.forward returns in the shipped setting (spatial, calculate_n_clusters "best",
    precompute_layouts.py:130-131) with np.random.seed(0) before every page batch (sklearn draws from the global state).
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.ref_import import import_reference  # noqa: E402
from rag_docvqa_b200 import synth  # noqa: E402

CASES = [dict(seed=11, pages=6, max_layouts=8, max_words=120), dict(seed=12, pages=8, max_layouts=30, max_words=300),
         dict(seed=13, pages=5, max_layouts=14, max_words=200)]


def hexlist(a: np.ndarray):
    return [float(x).hex() for x in np.asarray(a, dtype=np.float64).reshape(-1)]


def main():
    modules, _, _ = import_reference()
    emb = synth.HashEmbedder(384)
    emb.bge_model = type("M", (), {"tokenizer": None})()
    out = []
    for case in CASES:
        layout_info, pages_info = synth.make_s2_pages(**case)
        rec = dict(case=case, pages=[])
        for mode in ("spatial", "spatial+semantic"):
            s2 = modules.S2Chunker({"cluster_mode": mode, "calculate_n_clusters": "best"}, embedder=emb)
            for p, page in enumerate(layout_info):
                if len(page["boxes"]) == 0:
                    continue
                nodes, edges, used = s2.create_nodes_and_edges(page, pages_info[p] if mode != "spatial" else None)
                item = dict(mode=mode, page=p, ids=[n["global_id"] for n in nodes], used=[bool(u) for u in used],
                            texts_crc=[__import__("zlib").crc32(n["text"].encode()) for n in nodes], n_edges=len(edges))
                if len(nodes) >= 1:
                    item["weights"] = hexlist(s2._combined_weights(nodes))
                rec["pages"].append(item)
        s2 = modules.S2Chunker({"cluster_mode": "spatial", "calculate_n_clusters": "best"}, embedder=emb)
        np.random.seed(0)
        rec["clusters_spatial_best"] = [np.asarray(c).astype(int).tolist() for c in s2.forward(layout_info)]
        out.append(rec)
    path = os.path.join(ROOT, "tests", "golden", "s2chunker.json")
    with open(path, "w") as f:
        json.dump(out, f)
    print("wrote", path, os.path.getsize(path), "bytes", [len(r["pages"]) for r in out])


if __name__ == "__main__":
    main()
