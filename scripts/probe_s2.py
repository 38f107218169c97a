"""Timings of the S2Chunker drop-in (second half): the weight matrices of a batch of pages in one launch,
the batched node building, and forward() end to end, each beside the oracle (the reference's Python / numpy loops).
    python scripts/probe_s2.py            (on a B200; prints one JSON object)"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from oracle import ref_restated as R
from rag_docvqa_b200 import _lib, synth
from rag_docvqa_b200.functional import _stream_ptr
from rag_docvqa_b200.s2chunker import S2Chunker

dev = torch.device("cuda:0")


def best(fn, reps=5):
    out = 1e9
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        out = min(out, time.perf_counter() - t0)
    return out


# precompute_layouts.py:117-118 works on batches of 24 page images; a page has up to ~30 layout regions
layout_info, pages_info = synth.make_s2_pages(seed=5, pages=24, max_layouts=30, max_words=400, degenerate=False)
boxes = [p["boxes"] for p in layout_info]
entries = int(sum(len(b) ** 2 for b in boxes))
emb = synth.HashEmbedder(384, device=dev)
s2 = S2Chunker({"cluster_mode": "spatial", "calculate_n_clusters": "best", "device": "cuda:0"})
s2sem = S2Chunker({"cluster_mode": "spatial+semantic", "calculate_n_clusters": "best", "device": "cuda:0"}, embedder=emb)
embs = [torch.randn(len(b), 384, generator=torch.Generator().manual_seed(i)).to(dev) for i, b in enumerate(boxes)]
s2.weights_batch(boxes)
s2.weights_batch(boxes, embs)
out = {"pages": len(boxes), "regions": int(sum(len(b) for b in boxes)), "matrix_entries": entries}
out["weights_spatial_ms_device_call"] = 1e3 * best(lambda: s2.weights_batch(boxes))                # upload + launch + read-back
out["weights_combined_384d_ms_device_call"] = 1e3 * best(lambda: s2.weights_batch(boxes, embs))
out["weights_spatial_ms_oracle"] = 1e3 * best(lambda: [R.s2_spatial_weights(b) for b in boxes], reps=2)
embs_h = [e.cpu().numpy() for e in embs]
out["weights_combined_384d_ms_oracle"] = 1e3 * best(lambda: [R.s2_combined_weights(b, e) for b, e in zip(boxes, embs_h)], reps=2)

# the kernel alone (device-resident inputs, CUDA events)
n = np.asarray([len(b) for b in boxes], dtype=np.int64)
node_off = torch.from_numpy(np.concatenate([[0], np.cumsum(n)]).astype(np.int32)).to(dev)
out_off = torch.from_numpy(np.concatenate([[0], np.cumsum(n * n)]).astype(np.int64)).to(dev)
box = torch.from_numpy(np.concatenate([np.asarray(b, dtype=np.float64).reshape(-1, 4) for b in boxes])).to(dev)
e_all = torch.cat(embs).contiguous()
w = torch.empty(entries, dtype=torch.float64, device=dev)
for name, e_ptr, what in (("spatial", None, _lib.S2_SPATIAL), ("combined_384d", e_all.data_ptr(), _lib.S2_COMBINED)):
    def launch():
        _lib.check(_lib.lib.rdv_s2_weights(box.data_ptr(), node_off.data_ptr(), len(boxes), e_ptr, 384, what, out_off.data_ptr(),
                                           entries, w.data_ptr(), _stream_ptr(dev)))
    for _ in range(5):
        launch()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(100):
        launch()
    e1.record()
    torch.cuda.synchronize()
    out["kernel_%s_us" % name] = 1e3 * e0.elapsed_time(e1) / 100

# node building with the word -> region membership (one rdv_layout_assign launch for the batch) vs the oracle's loops
out["word_region_pairs"] = int(sum(len(p["boxes"]) * len(q["ocr_tokens"]) for p, q in zip(layout_info, pages_info)))
s2sem._nodes_batch(layout_info, pages_info)
out["nodes_semantic_ms_device_call"] = 1e3 * best(lambda: s2sem._nodes_batch(layout_info, pages_info))
out["nodes_semantic_ms_oracle"] = 1e3 * best(lambda: [R.s2_nodes(p, q, "spatial+semantic") for p, q in zip(layout_info, pages_info)], reps=2)

# forward() in the shipped setting (spatial / best): the sklearn clustering dominates both arms
s2.forward(layout_info[:3]); R.s2_forward(layout_info[:3], None, "spatial")      # first-use imports (sklearn, networkx) out of the timing
np.random.seed(0)
t0 = time.perf_counter(); got = s2.forward(layout_info); out["forward_spatial_best_ms"] = 1e3 * (time.perf_counter() - t0)
np.random.seed(0)
t0 = time.perf_counter(); ref = R.s2_forward(layout_info, None, "spatial"); out["forward_spatial_best_ms_oracle"] = 1e3 * (time.perf_counter() - t0)
out["forward_equal"] = [np.asarray(a).tolist() for a in got] == [np.asarray(a).tolist() for a in ref]
print(json.dumps(out))
