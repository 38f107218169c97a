"""Wall-clock phases of the pipelined host path of Retriever.retrieve on C2 (where do the 2.6 ms go?)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rag_docvqa_b200 import synth, functional as F
from rag_docvqa_b200.retriever import Retriever
dev = torch.device("cuda:0")
b = synth.make_text_batch("C2", with_lists=True, share_image_pool=24)
emb = [e.pin_memory() for e in b["text_embeddings"]]
q = b["question_embeddings"].pin_memory()
lists = (b["words_text_chunks"], b["words_box_chunks"], b["layout_labels_chunks"], b["images"], b["page_indices"])
r = Retriever({"compute_stats": False, "compute_stats_examples": False, "n_stats_examples": 0, "chunk_num": 5,
               "device": "cuda:0", "retrieval_lazy_patches": True})
for _ in range(5):
    r.retrieve(emb, q, *lists)
torch.cuda.synchronize()
T = time.perf_counter
# 1. pure H2D of the batch
t0 = T()
for _ in range(10):
    tab = F.upload_doc_table(emb, 384, dev)
    torch.cuda.synchronize()
print("upload_doc_table + sync: %.3f ms" % ((T() - t0) / 10 * 1e3))
t0 = T()
for _ in range(10):
    tab = F.upload_doc_table(emb, 384, dev)
t1 = T()
torch.cuda.synchronize()
print("upload_doc_table enqueue only: %.3f ms" % ((t1 - t0) / 10 * 1e3))
# 2. one big copy for comparison
big = torch.cat(emb).pin_memory()
t0 = T()
for _ in range(10):
    big.to(dev, non_blocking=True)
    torch.cuda.synchronize()
print("single 32 MB pinned copy + sync: %.3f ms (%.1f GB/s)" % ((T() - t0) / 10 * 1e3, big.numel() * 4 / ((T() - t0) / 10) / 1e9))
# 3. list building alone
res = r._score_topk(emb, q)
hits = r._hits_to_host(res.topk_idx, res.topk_cnt)
t0 = T()
for _ in range(10):
    r._hit_lists(hits, *lists)
print("_hit_lists alone: %.3f ms" % ((T() - t0) / 10 * 1e3))
import gc
gc.disable()
t0 = T()
for _ in range(10):
    r._hit_lists(hits, *lists)
print("_hit_lists alone, gc disabled: %.3f ms" % ((T() - t0) / 10 * 1e3))
t0 = T()
for _ in range(20):
    r.retrieve(emb, q, *lists)
torch.cuda.synchronize()
print("retrieve, gc disabled: %.3f ms" % ((T() - t0) / 20 * 1e3))
gc.enable()
t0 = T()
for _ in range(20):
    out = r.retrieve(emb, q, *lists)
torch.cuda.synchronize()
print("retrieve: %.3f ms" % ((T() - t0) / 20 * 1e3))
t0 = T()
for _ in range(20):
    out = r.retrieve(emb, q, *lists)
    del out
torch.cuda.synchronize()
print("retrieve (result dropped each time): %.3f ms" % ((T() - t0) / 20 * 1e3))
# 4. zero-copy scoring: the kernel streams the pinned rows over PCIe
tab = F.upload_doc_table(emb, 384, dev)
qd = q.to(dev)
for _ in range(3):
    F.score_topk_table(tab, qd, 5)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    F.score_topk_table(tab, qd, 5)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print("zero-copy score+topk of the pinned batch: %.3f ms (%.1f GB/s over PCIe)" % (ms, big.numel() * 4 / ms / 1e6))
for tr in (8, 16, 64, 128):
    tab2 = F.upload_doc_table(emb, 384, dev, tile_rows=tr)
    F.score_topk_table(tab2, qd, 5)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(10):
        F.score_topk_table(tab2, qd, 5)
    e1.record()
    torch.cuda.synchronize()
    print("  tile_rows=%d: %.3f ms" % (tr, e0.elapsed_time(e1) / 10))
