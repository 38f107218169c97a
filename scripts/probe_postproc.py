"""Timings of the rank-3 / rank-4 kernels: rerank order, page vote, reranked re-emission, layout assignment.
    python scripts/probe_postproc.py            (on a B200; prints one JSON object)"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from rag_docvqa_b200 import postproc, synth
from rag_docvqa_b200.chunker import Chunker

dev = torch.device("cuda:0")


def timed(fn, reps=50, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


out = {}
rng = np.random.RandomState(0)
for name, B, k, n_b in (("C2", 64, 5, 600), ("C3", 256, 10, 10000)):
    scores = torch.from_numpy(rng.rand(B, k).astype(np.float32)).to(dev)
    cnt = torch.full((B,), k, dtype=torch.int32, device=dev)
    pages = torch.from_numpy(rng.randint(0, 20, size=(B, k)).astype(np.int32)).to(dev)
    sims = torch.from_numpy((0.2 + 0.4 * rng.rand(B * n_b)).astype(np.float32)).to(dev)      # cosines cluster in 0.2 .. 0.6
    row_off = torch.arange(0, (B + 1) * n_b, n_b, dtype=torch.int64, device=dev)
    out[name] = {
        "rerank_order_us": 1e3 * timed(lambda: postproc.rerank_order(scores, cnt, 0.4, 5, 1)),
        "page_vote_major_us": 1e3 * timed(lambda: postproc.page_vote(pages, cnt, None, row_off, False)),
        "page_vote_weighted_us": 1e3 * timed(lambda: postproc.page_vote(pages, cnt, sims, row_off, True)),
        "page_vote_weighted_f32_accumulation_us": 1e3 * timed(lambda: postproc.page_vote(pages, cnt, sims, row_off, True, legacy_promotion=False)),
    }

# layout assignment: a C2-shaped batch of pages (64 documents x <= 20 pages, <= 700 words, <= 30 layout boxes)
words, boxes, info = synth.make_chunker_batch(77, docs=64, max_pages=20, max_words=700, max_layouts=30, degenerate=False)
cfg = {"compute_stats": False, "compute_stats_examples": False, "n_stats_examples": 0, "device": "cuda:0"}
ch = Chunker(cfg)
pb = [np.asarray(p, dtype=np.float64).reshape(-1, 4) for d in boxes for p in d]
lb = [np.asarray(pg["boxes"], dtype=np.float64).reshape(-1, 4) for d in info for pg in d]
ll = [np.arange(len(x), dtype=np.int32) for x in lb]
pairs = int(sum(len(a) * len(b) for a, b in zip(pb, lb)))
t0 = time.perf_counter(); ch.assign_words_to_layouts(pb, lb, ll); torch.cuda.synchronize(); t1 = time.perf_counter()
reps = 5
t0 = time.perf_counter()
for _ in range(reps):
    ch.assign_words_to_layouts(pb, lb, ll)
torch.cuda.synchronize()
assign_ms = (time.perf_counter() - t0) / reps * 1e3
t0 = time.perf_counter()
res = ch.get_chunks(words, boxes, info, question_id=list(range(len(words))))
gpu_s = time.perf_counter() - t0
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_restated as R
t0 = time.perf_counter()
want, _ = R.get_chunks(words, boxes, info)
cpu_s = time.perf_counter() - t0
out["chunker"] = {"pages": len(pb), "words": int(sum(len(a) for a in pb)), "word_x_box_pairs": pairs,
                  "assign_ms_incl_upload_and_readback": assign_ms, "get_chunks_s": gpu_s, "oracle_get_chunks_s": cpu_s,
                  "chunks": int(sum(len(x) for x in res[0])), "equal": json.loads(json.dumps(res)) == json.loads(json.dumps(want))}
print(json.dumps(out))
