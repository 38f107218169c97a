"""A few launches of the packed-cluster kernel (score + top-k) and of the streaming tile kernel on C2 batches: the
target of ncu captures.   python scripts/probe_cluster.py [n_launches]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rag_docvqa_b200 import functional as F, synth, _lib

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
dev = torch.device("cuda:0")
torch.cuda.set_device(0)
w = synth.WORKLOADS["C2"]
R = 10
batches = [synth.make_text_batch("C2", device=dev, seed=synth.SEED_BASE + 2, emb_seed=1000 * (r + 1)) for r in range(R)]
tables = [F.build_doc_table(b["text_embeddings"], w.dim, dev) for b in batches]
t = tables[0]
print("tiles %d, cluster CTAs %d (clusters of %d)" % (t.total_tiles, t.n_ctas, tables[0].cluster))
sims = torch.empty(t.total_rows, device=dev)
idx = torch.empty((t.B, w.k), dtype=torch.int32, device=dev)
val = torch.empty((t.B, w.k), device=dev)
cnt = torch.empty((t.B,), dtype=torch.int32, device=dev)
s = torch.cuda.current_stream().cuda_stream
for i in range(n):
    t, b = tables[i % R], batches[i % R]
    d_ctas, n_ctas, cl = t.cluster_pointers()
    _lib.check(_lib.lib.rdv_score_topk_cluster_f32(d_ctas, n_ctas, cl, b["question_embeddings"].data_ptr(), t.B, t.d, w.k, t.max_rows,
                                                   sims.data_ptr(), idx.data_ptr(), val.data_ptr(), cnt.data_ptr(), s))
    _lib.check(_lib.lib.rdv_score_f32(t.pointers()[0], t.total_tiles, t.tile_rows, _lib.SCORE_LDG, b["question_embeddings"].data_ptr(),
                                      t.B, t.d, sims.data_ptr(), s))
torch.cuda.synchronize()
print("ok")
