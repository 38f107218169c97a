"""TEST INFRASTRUCTURE ONLY -- freezes the reference's Chunker.get_chunks (src/_modules.py:843-1100).

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden_chunker.py

Runs the UNMODIFIED reference Chunker (needs /root/reference) on seeded synthetic pages (rag_docvqa_b200.synth.
make_chunker_batch: inputs are regenerated from the seed by the tests) and writes tests/golden/chunker.json: for small
cases the full 5-tuple, for larger ones a CRC of its JSON form, plus the stats counters.
"""
import json
import os
import sys
import zlib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.ref_import import import_reference  # noqa: E402
from rag_docvqa_b200 import synth  # noqa: E402

CASES = [
    # seed, docs, max_pages, max_words, max_layouts, clusters, cluster_layouts, chunk_size, overlap, tol, page_retrieval, full
    dict(seed=1, docs=2, max_pages=2, max_words=90, max_layouts=5, clusters=False, cluster_layouts=False, chunk_size=20, overlap=5, tol=0.2, page_retrieval="concat", full=True),
    dict(seed=2, docs=2, max_pages=2, max_words=90, max_layouts=5, clusters=True, cluster_layouts=True, chunk_size=20, overlap=5, tol=0.2, page_retrieval="concat", full=True),
    dict(seed=3, docs=4, max_pages=6, max_words=400, max_layouts=14, clusters=False, cluster_layouts=False, chunk_size=60, overlap=10, tol=0.2, page_retrieval="concat", full=False),
    dict(seed=4, docs=4, max_pages=6, max_words=400, max_layouts=14, clusters=True, cluster_layouts=True, chunk_size=60, overlap=10, tol=0.2, page_retrieval="concat", full=False),
    dict(seed=5, docs=3, max_pages=5, max_words=300, max_layouts=10, clusters=True, cluster_layouts=False, chunk_size=30, overlap=0, tol=0.0, page_retrieval="concat", full=False),
    dict(seed=6, docs=3, max_pages=5, max_words=300, max_layouts=10, clusters=False, cluster_layouts=False, chunk_size=7, overlap=6, tol=1.0, page_retrieval="maxconf", full=False),
    dict(seed=7, docs=3, max_pages=4, max_words=200, max_layouts=8, clusters=False, cluster_layouts=False, chunk_size=60, overlap=10, tol=0.2, page_retrieval="oracle", full=False),
    dict(seed=8, docs=3, max_pages=4, max_words=200, max_layouts=8, clusters=False, cluster_layouts=False, chunk_size=60, overlap=10, tol=0.2, page_retrieval="concat", full=False, no_layout=True),
    dict(seed=9, docs=3, max_pages=4, max_words=200, max_layouts=8, clusters=False, cluster_layouts=False, chunk_size=60, overlap=10, tol=0.2, page_retrieval="concat", full=False, numpy_pages=True),
]


def crc(obj) -> int:
    return zlib.crc32(json.dumps(obj).encode()) & 0xFFFFFFFF


def config_of(case):
    return {"compute_stats": True, "compute_stats_examples": False, "n_stats_examples": 0, "layout_model_weights": None,
            "chunk_size": case["chunk_size"], "overlap": case["overlap"], "chunk_size_tol": case["tol"],
            "page_retrieval": case["page_retrieval"], "cluster_layouts": case["cluster_layouts"]}


def inputs_of(case):
    words, boxes, info = synth.make_chunker_batch(case["seed"], case["docs"], case["max_pages"], case["max_words"],
                                                  case["max_layouts"], clusters=case["clusters"],
                                                  numpy_pages=case.get("numpy_pages", False))
    return words, boxes, ([[]] if case.get("no_layout") else info)


def stats_json(stats):
    return {k: {str(a): int(b) for a, b in v.items()} for k, v in stats.items()}


def main():
    modules, _, _ = import_reference()
    out = []
    for case in CASES:
        words, boxes, info = inputs_of(case)
        ch = modules.Chunker(config_of(case))
        res = ch.get_chunks(words, boxes, info, question_id=["q%d" % b for b in range(len(words))])
        res = json.loads(json.dumps(res))
        rec = dict(case=case, crc=[crc(x) for x in res], stats=stats_json(ch.stats),
                   n_chunks=[len(x) for x in res[0]])
        if case["full"]:
            rec["outputs"] = res
        out.append(rec)
    path = os.path.join(ROOT, "tests", "golden", "chunker.json")
    with open(path, "w") as f:
        json.dump(out, f)
    print("wrote", path, os.path.getsize(path), "bytes", [r["n_chunks"] for r in out])


if __name__ == "__main__":
    main()
