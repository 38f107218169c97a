"""Per-CTA timeline of the cluster kernels on a C2 batch (rdv_debug_trace): where do the microseconds go?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from rag_docvqa_b200 import functional as F, synth, _lib
from rag_docvqa_b200.docstore import DocStore
import bench

dev = torch.device("cuda:0")
torch.cuda.set_device(0)
w = synth.WORKLOADS["C2"]
host_batch = synth.make_text_batch("C2", with_lists=True, share_image_pool=24)
R = 10
batches = [synth.make_text_batch("C2", device=dev, seed=synth.SEED_BASE + 2, emb_seed=1000 * (r + 1)) for r in range(R)]
tables = [F.build_doc_table(b["text_embeddings"], w.dim, dev) for b in batches]
t0 = tables[0]
print("cluster table: %d CTAs in clusters of %d" % (t0.n_ctas, t0.cluster))
table_w = synth.make_tokens_for_words(host_batch["words_text_chunks"], seed=3)
store = DocStore.from_lists(host_batch["words_text_chunks"], host_batch["words_box_chunks"], host_batch["layout_labels_chunks"],
                            host_batch["page_indices"], lambda wd: table_w.get(wd, [2]), dev, images=host_batch["images"])
prompts = bench.prompts_for(w.docs)
sims = torch.empty(t0.total_rows, device=dev)
idx = torch.empty((t0.B, w.k), dtype=torch.int32, device=dev)
val = torch.empty((t0.B, w.k), device=dev)
cnt = torch.empty((t0.B,), dtype=torch.int32, device=dev)
plan = store.prepare_gather(idx, cnt, prompts, max_len=512, sims=sims, topk_val=val, max_rows=t0.max_rows)
trace = torch.zeros((t0.n_ctas, 8), dtype=torch.int64, device=dev)
lib = _lib.lib
s = torch.cuda.current_stream().cuda_stream
ctas = np.frombuffer(t0.desc.cpu().numpy().tobytes()[t0.ctas_offset:t0.ctas_offset + 32 * t0.n_ctas], dtype=F.CTA_DTYPE)
for mode in ("topk", "retrieve"):
    rows = []
    for i in range(12):
        t, b = tables[i % R], batches[i % R]
        d_ctas, n_ctas, cl = t.cluster_pointers()
        trace.zero_()
        lib.rdv_debug_trace(trace.data_ptr() if i >= 6 else None)
        if mode == "topk":
            _lib.check(lib.rdv_score_topk_cluster_f32(d_ctas, n_ctas, cl, b["question_embeddings"].data_ptr(), t.B, t.d, w.k, t.max_rows,
                                                      sims.data_ptr(), idx.data_ptr(), val.data_ptr(), cnt.data_ptr(), s))
        else:
            _lib.check(lib.rdv_retrieve_vt5_f32(d_ctas, n_ctas, cl, b["question_embeddings"].data_ptr(), t.d, t.max_rows, sims.data_ptr(),
                                                plan._ds_ref, plan._args_ref, s))
        torch.cuda.synchronize()
        if i >= 6:
            rows.append(trace.cpu().numpy().astype(np.float64))
    lib.rdv_debug_trace(None)
    tr = rows[-1]
    t_first = tr[:, 0][tr[:, 0] > 0].min()
    rel = np.where(tr > 0, (tr - t_first) / 1e3, np.nan)          # us since the first CTA started
    names = ["started", "rows streamed", "cluster complete", "candidates pushed", "candidates here", "merged", "gathered"]
    print("== %s: us since the first CTA started (min / median / max over CTAs that reach the point)" % mode)
    for j, nm in enumerate(names):
        col = rel[:, j][~np.isnan(rel[:, j])]
        if len(col):
            print("  %-18s n=%4d  %6.2f / %6.2f / %6.2f" % (nm, len(col), col.min(), np.median(col), col.max()))
    cl_id = np.arange(t0.n_ctas) // t0.cluster
    start_by_cluster = np.array([np.nanmin(rel[cl_id == c, 0]) for c in range(t0.n_ctas // t0.cluster)])
    end_by_cluster = np.array([np.nanmax(rel[cl_id == c, :]) for c in range(t0.n_ctas // t0.cluster)])
    print("  cluster start times:", np.round(np.sort(start_by_cluster), 2).tolist())
    print("  cluster end times:  ", np.round(np.sort(end_by_cluster), 2).tolist())
    dur = rel[:, 1] - rel[:, 0]
    print("  streaming duration per CTA: median %.2f max %.2f (rows %s)" % (np.nanmedian(dur), np.nanmax(dur), ""))
