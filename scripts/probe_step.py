"""Where does the C2 step time go: host enqueue cost vs device time; plain launches vs CUDA-graph replay."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rag_docvqa_b200 import functional as F, synth, _lib
from rag_docvqa_b200.docstore import DocStore
import bench

dev = torch.device("cuda:0")
torch.cuda.set_device(0)
w = synth.WORKLOADS["C2"]
host_batch = synth.make_text_batch("C2", with_lists=True, share_image_pool=24)
R = 9
batches = [synth.make_text_batch("C2", device=dev, seed=synth.SEED_BASE + 2, emb_seed=1000 * (r + 1)) for r in range(R)]
tables = [F.build_doc_table(b["text_embeddings"], w.dim, dev) for b in batches]
outs = [dict(sims=torch.empty(t.total_rows, dtype=torch.float32, device=dev), idx=torch.empty((t.B, w.k), dtype=torch.int32, device=dev),
             val=torch.empty((t.B, w.k), dtype=torch.float32, device=dev), cnt=torch.empty((t.B,), dtype=torch.int32, device=dev)) for t in tables]
table = synth.make_tokens_for_words(host_batch["words_text_chunks"], seed=3)
store = DocStore.from_lists(host_batch["words_text_chunks"], host_batch["words_box_chunks"], host_batch["layout_labels_chunks"],
                            host_batch["page_indices"], lambda wd: table.get(wd, [2]), dev, images=host_batch["images"])
prompts = bench.prompts_for(w.docs)
plans = [store.prepare_gather(o["idx"], o["cnt"], prompts, max_len=512, sims=o["sims"], topk_val=o["val"], max_rows=t.max_rows)
         for o, t in zip(outs, tables)]
lib = _lib.lib

def make_step(stream):
    sa = []
    for t, o, b in zip(tables, outs, batches):
        p_tiles, _ = t.pointers()
        sa.append((p_tiles, t.total_tiles, t.tile_rows, t.algo, b["question_embeddings"].data_ptr(), t.B, t.d, o["sims"].data_ptr(), stream))
    def score(i): lib.rdv_score_f32(*sa[i % R])
    def gather(i): plans[i % R].launch(stream)
    def step(i): score(i); gather(i)
    return score, gather, step

def timed(fn, n):
    for i in range(20): fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for i in range(n): fn(i)
    t_host = time.perf_counter() - t0
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3, t_host / n * 1e6

score, gather, step = make_step(torch.cuda.current_stream().cuda_stream)
for name, fn in (("score", score), ("gather", gather), ("step", step)):
    dev_us, host_us = timed(fn, 900)
    print("%-8s plain launches: device %.2f us, host enqueue %.2f us per call" % (name, dev_us, host_us))

# CUDA graph: R steps per replay
side = torch.cuda.Stream()
for name in ("score", "gather", "step"):
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        sc, ga, st = make_step(side.cuda_stream)
        fn = dict(score=sc, gather=ga, step=st)[name]
        for i in range(R): fn(i)
        side.synchronize()
        with torch.cuda.graph(g, stream=side):
            for i in range(R): fn(i)
    for _ in range(5): g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(100): g.replay()
    e1.record()
    torch.cuda.synchronize()
    print("%-8s graph replay (%d per graph): device %.2f us per call" % (name, R, e0.elapsed_time(e1) / 100 / R * 1e3))
