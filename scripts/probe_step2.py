"""C2 / C3 step as ONE dependent chain: two launches (score + select/gather) against the one-launch kernel
(rdv_retrieve_vt5_f32) and the one-launch score + top-k (rdv_score_topk_cluster_f32), plain launches and CUDA-graph replay.
    python scripts/probe_step2.py [C2|C3] [tile_rows]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from rag_docvqa_b200 import functional as F, synth, _lib
from rag_docvqa_b200.docstore import DocStore
import bench

wl = sys.argv[1] if len(sys.argv) > 1 else "C2"
tile_rows = int(sys.argv[2]) if len(sys.argv) > 2 else 0
dev = torch.device("cuda:0")
torch.cuda.set_device(0)
w = synth.WORKLOADS[wl]
with_lists = wl != "C3"
host_batch = synth.make_text_batch(wl, with_lists=with_lists, share_image_pool=24)
step_bytes = bench.score_bytes(host_batch["sizes"], w.dim, w.k)
R = max(3, int(np.ceil(2.5 * bench.L2_BYTES / step_bytes)))
batches = [synth.make_text_batch(wl, device=dev, seed=synth.SEED_BASE + w.config_id, emb_seed=1000 * (r + 1)) for r in range(R)]
tables = [F.build_doc_table(b["text_embeddings"], w.dim, dev, tile_rows=tile_rows) for b in batches]
print("%s: %d docs, %d rows, %.1f MB/step, R=%d, tile_rows=%d, tiles=%d" % (wl, tables[0].B, tables[0].total_rows, step_bytes / 1e6, R,
                                                                          tables[0].tile_rows, tables[0].total_tiles))
outs = [dict(sims=torch.empty(t.total_rows, dtype=torch.float32, device=dev), idx=torch.empty((t.B, w.k), dtype=torch.int32, device=dev),
             val=torch.empty((t.B, w.k), dtype=torch.float32, device=dev), cnt=torch.empty((t.B,), dtype=torch.int32, device=dev))
        for t in tables]
cluster_ok = tables[0].n_ctas > 0
print("cluster table: %d CTAs in clusters of %d" % (tables[0].n_ctas, tables[0].cluster))
plans = None
if with_lists:
    table = synth.make_tokens_for_words(host_batch["words_text_chunks"], seed=3)
    store = DocStore.from_lists(host_batch["words_text_chunks"], host_batch["words_box_chunks"], host_batch["layout_labels_chunks"],
                                host_batch["page_indices"], lambda wd: table.get(wd, [2]), dev, images=host_batch["images"])
    prompts = bench.prompts_for(w.docs)
    plans = [store.prepare_gather(o["idx"], o["cnt"], prompts, max_len=512, sims=o["sims"], topk_val=o["val"], max_rows=t.max_rows)
             for o, t in zip(outs, tables)]
lib = _lib.lib


def make(stream):
    qs = [b["question_embeddings"] for b in batches]

    def score(i):
        t, o = tables[i % R], outs[i % R]
        _lib.check(lib.rdv_score_f32(t.pointers()[0], t.total_tiles, t.tile_rows, _lib.SCORE_LDG, qs[i % R].data_ptr(), t.B, t.d,
                                     o["sims"].data_ptr(), stream))

    def select(i):
        t, o = tables[i % R], outs[i % R]
        if plans is not None:
            plans[i % R].launch(stream)
        else:
            _lib.check(lib.rdv_topk_segments_f32(o["sims"].data_ptr(), t.pointers()[1], t.B, w.k, t.max_rows, o["idx"].data_ptr(),
                                                 o["val"].data_ptr(), o["cnt"].data_ptr(), stream))

    def two(i):
        score(i); select(i)

    def topk1(i):
        t, o = tables[i % R], outs[i % R]
        d_ctas, n_ctas, cl = t.cluster_pointers()
        _lib.check(lib.rdv_score_topk_cluster_f32(d_ctas, n_ctas, cl, qs[i % R].data_ptr(), t.B, t.d, w.k, t.max_rows,
                                                  o["sims"].data_ptr(), o["idx"].data_ptr(), o["val"].data_ptr(), o["cnt"].data_ptr(),
                                                  stream))

    def one(i):
        t = tables[i % R]
        d_ctas, n_ctas, cl = t.cluster_pointers()
        _lib.check(lib.rdv_retrieve_vt5_f32(d_ctas, n_ctas, cl, qs[i % R].data_ptr(), t.d, t.max_rows,
                                            outs[i % R]["sims"].data_ptr(), plans[i % R]._ds_ref, plans[i % R]._args_ref, stream))

    def topk1_gather(i):
        topk1(i)
        a = plans[i % R].args
        keep = a.sims
        a.sims = None
        plans[i % R].launch(stream)
        a.sims = keep
    fns = {"score": score, "select(+gather)": select, "two launches": two}
    if cluster_ok:
        fns["score+topk, one launch (cluster)"] = topk1
        if plans is not None:
            fns["retrieve, one launch (cluster)"] = one
            fns["score+topk one launch, then gather"] = topk1_gather
    return fns


def timed_plain(fn, n):
    for i in range(20): fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n): fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


n_plain = 400 if wl != "C3" else 40
K = 2 * R if wl != "C3" else R
side = torch.cuda.Stream()
for name, fn in make(torch.cuda.current_stream().cuda_stream).items():
    us_plain = timed_plain(fn, n_plain)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        gfn = make(side.cuda_stream)[name]
        for i in range(R): gfn(i)
        side.synchronize()
        with torch.cuda.graph(g, stream=side):
            for i in range(K): gfn(i)
    for _ in range(5): g.replay()
    torch.cuda.synchronize()
    reps = 100 if wl != "C3" else 10
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): g.replay()
    e1.record()
    torch.cuda.synchronize()
    us_graph = e0.elapsed_time(e1) / reps / K * 1e3
    print("%-38s plain %8.2f us   graph (one chain) %8.2f us   = %.0f GB/s of step bytes" % (name, us_plain, us_graph, step_bytes / us_graph / 1e3))
