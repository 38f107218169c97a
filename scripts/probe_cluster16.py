"""Does the device take clusters of 16 CTAs for the retrieval kernel?  (RDV_PDL=0 / 1, a few shapes.)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rag_docvqa_b200 import functional as F, synth, _lib
dev = torch.device("cuda:0")
torch.cuda.set_device(0)
lib = _lib.lib
for d in (128, 384, 768, 100):
    for k in (5, 10, 20):
        pass
for sizes in ([30], [600, 30, 0, 100]):
    emb, q = synth.make_embeddings(sizes, 384, 3)
    emb = [e.to(dev) for e in emb]; q = q.to(dev)
    t = F.build_doc_table(emb, 384, dev)
    print("sizes", sizes, "table: n_ctas", t.n_ctas, "cluster", t.cluster)
    try:
        r = F.score_topk_table(t, q, 5, cluster=True)
        torch.cuda.synchronize()
        print("  launch ok:", r.topk_idx.cpu().tolist()[0])
    except Exception as e:
        print("  launch failed:", e)
        torch.cuda.synchronize()
