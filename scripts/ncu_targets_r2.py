"""Two launches each of the kernels whose ncu summaries were missing after round 1 plus the C2 / C3
streaming kernel, the gather kernel and the cluster kernel -- the target of scripts/gpu_run_r2_ncu.sh.
    python scripts/ncu_targets_r2.py [all|score]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from rag_docvqa_b200 import _lib, functional as F, sharded, synth
from rag_docvqa_b200.docstore import DocStore

what = sys.argv[1] if len(sys.argv) > 1 else "all"
dev = torch.device("cuda:0")
torch.cuda.set_device(0)
lib = _lib.lib
s = torch.cuda.current_stream().cuda_stream
REPS = 2

# ---- C2: streaming score kernel (LDG and TMA), stand-alone selection, select + gather, cluster kernel ------------------
w = synth.WORKLOADS["C2"]
host = synth.make_text_batch("C2", with_lists=True, share_image_pool=24)
b2 = [synth.make_text_batch("C2", device=dev, seed=synth.SEED_BASE + 2, emb_seed=1000 * (r + 1)) for r in range(REPS)]
for algo in (_lib.SCORE_LDG, _lib.SCORE_TMA):
    for b in b2:
        t = F.build_doc_table(b["text_embeddings"], w.dim, dev, algo=algo)
        sims = F.score_table(t, b["question_embeddings"])
if what == "all":
    for b in b2:
        t = F.build_doc_table(b["text_embeddings"], w.dim, dev)
        res = F.score_topk_table(t, b["question_embeddings"], w.k, cluster=False)          # topk_segments_kernel
    table_w = synth.make_tokens_for_words(host["words_text_chunks"], seed=3)
    store = DocStore.from_lists(host["words_text_chunks"], host["words_box_chunks"], host["layout_labels_chunks"],
                                host["page_indices"], lambda wd: table_w.get(wd, [2]), dev, images=host["images"])
    prompts = [[5, 6, 7, 8 + i] for i in range(w.docs)]
    for b in b2:
        t = F.build_doc_table(b["text_embeddings"], w.dim, dev)
        sims = torch.empty(t.total_rows, device=dev)
        idx = torch.empty((t.B, w.k), dtype=torch.int32, device=dev)
        val = torch.empty((t.B, w.k), device=dev)
        cnt = torch.empty((t.B,), dtype=torch.int32, device=dev)
        plan = store.prepare_gather(idx, cnt, prompts, max_len=512, sims=sims, topk_val=val, max_rows=t.max_rows)
        F.score_table(t, b["question_embeddings"], out=sims)
        plan.launch()                                                                      # gather_vt5_kernel
        plan.launch_retrieve(t, F._f32_contig_aligned(b["question_embeddings"]), sims)      # retrieve_cluster_kernel
torch.cuda.synchronize()
del b2

# ---- C3: the streaming kernel on long documents ------------------------------------------------------------------------
w3 = synth.WORKLOADS["C3"]
b3 = synth.make_text_batch("C3", device=dev)
t3 = F.build_doc_table(b3["text_embeddings"], w3.dim, dev)
for _ in range(REPS):
    sims3 = F.score_table(t3, b3["question_embeddings"])
if what == "all":
    for _ in range(REPS):
        F.score_topk_table(t3, b3["question_embeddings"], w3.k)                            # topk_segments_kernel at 10 k rows
torch.cuda.synchronize()
del b3, t3, sims3
torch.cuda.empty_cache()

if what == "all":
    # ---- mean pooling: 8192 chunks x <= 160 tokens x 384 ----------------------------------------------------------------
    embs, mask = synth.make_token_batch(8192, 384, 7, device=dev, max_len=160)
    for _ in range(REPS):
        F.mean_pooling(embs, mask)
    del embs, mask
    # ---- MaxSim: strict fp32 (FFMA), bf16 tcgen05, on 8 strips of one C4 question -------------------------------------
    patches, q = synth.make_strip_batch(1, [8], 2048, 768, 3, device=dev)
    for _ in range(REPS):
        F.late_interaction(q[0:1], patches[0], mode="ffma")
        F.late_interaction_bf16(q[0:1], patches[0])
    del patches, q
    # ---- corpus mode at C5's per-rank shape + the merge kernels -------------------------------------------------------
    g = torch.Generator(device=dev)
    g.manual_seed(5)
    rows = torch.empty((1_250_000, 768), dtype=torch.bfloat16, device=dev)
    for a in range(0, rows.shape[0], 250_000):
        rows[a:a + 250_000] = torch.randn(250_000, 768, generator=g, device=dev).to(torch.bfloat16)
    shard = sharded.CorpusShard(rows)
    qs = torch.randn(1024, 768, generator=g, device=dev)
    searcher = sharded.CorpusSearcher(shard, 1024, 10, graph=False)
    for _ in range(REPS):
        searcher.search(qs)                                                                # tc_score_kernel, topk_merge_kernel
    recv = torch.stack([searcher.send.clone() for _ in range(8)]).contiguous()             # what an 8-rank all-gather delivers
    ov = torch.empty((1024, 10), device=dev)
    oi = torch.empty((1024, 10), dtype=torch.int64, device=dev)
    for _ in range(REPS):
        _lib.check(lib.rdv_topk_merge_parts(recv.data_ptr(), recv.data_ptr() + searcher.nv, 1024, 8, 10, searcher.send.numel() // 4,
                                            searcher.send.numel() // 8, 10, ov.data_ptr(), oi.data_ptr(), s))
    del rows, shard, searcher
    # ---- S2Chunker weight matrices: 24 pages (spatial + semantic) -------------------------------------------------------
    from rag_docvqa_b200.s2chunker import S2Chunker
    layout, _ = synth.make_s2_pages(seed=5, pages=24, max_layouts=30, max_words=400, degenerate=False)
    boxes = [p["boxes"] for p in layout]
    s2 = S2Chunker({"cluster_mode": "spatial+semantic", "calculate_n_clusters": "best", "device": "cuda:0"},
                   embedder=synth.HashEmbedder(384, device=dev))
    embs = [torch.randn(len(b), 384, generator=torch.Generator().manual_seed(i)).to(dev) for i, b in enumerate(boxes)]
    for _ in range(REPS):
        s2.weights_batch(boxes, embs)
torch.cuda.synchronize()
print("ok")
