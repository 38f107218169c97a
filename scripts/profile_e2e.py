"""cProfile of the drop-in Retriever.retrieve on C2 with host inputs (where does the e2e time go?)."""
import sys, os, time, cProfile, pstats
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rag_docvqa_b200 import synth
from rag_docvqa_b200.retriever import Retriever
dev = "cuda:0"
b = synth.make_text_batch("C2", with_lists=True, share_image_pool=24)
emb = [e.pin_memory() for e in b["text_embeddings"]]
q = b["question_embeddings"].pin_memory()
lists = (b["words_text_chunks"], b["words_box_chunks"], b["layout_labels_chunks"], b["images"], b["page_indices"])
r = Retriever({"compute_stats": False, "compute_stats_examples": False, "n_stats_examples": 0, "chunk_num": 5,
               "device": dev, "retrieval_lazy_patches": True})
for _ in range(3):
    r.retrieve(emb, q, *lists)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20):
    r.retrieve(emb, q, *lists)
torch.cuda.synchronize()
print("ms per retrieve: %.3f" % ((time.perf_counter() - t0) / 20 * 1e3))
pr = cProfile.Profile()
pr.enable()
for _ in range(20):
    r.retrieve(emb, q, *lists)
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
