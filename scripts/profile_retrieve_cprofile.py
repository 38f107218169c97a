"""cProfile of the drop-in Retriever.retrieve on the C2 batch, pinned host embeddings (where does the host time go?)."""
import cProfile, pstats, sys, os, time, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rag_docvqa_b200 import synth
from rag_docvqa_b200.retriever import Retriever
b = synth.make_text_batch("C2", with_lists=True, share_image_pool=24)
emb = [e.pin_memory() for e in b["text_embeddings"]]
q = b["question_embeddings"].pin_memory()
lists = (b["words_text_chunks"], b["words_box_chunks"], b["layout_labels_chunks"], b["images"], b["page_indices"])
r = Retriever({"compute_stats": False, "compute_stats_examples": False, "n_stats_examples": 0, "chunk_num": 5,
               "device": "cuda:0", "retrieval_lazy_patches": True})
for _ in range(10):
    r.retrieve(emb, q, *lists)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(50):
    r.retrieve(emb, q, *lists)
torch.cuda.synchronize()
print("retrieve: %.3f ms per batch" % ((time.perf_counter() - t0) / 50 * 1e3))
pr = cProfile.Profile()
pr.enable()
for _ in range(50):
    r.retrieve(emb, q, *lists)
torch.cuda.synchronize()
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(28)
print(s.getvalue())
