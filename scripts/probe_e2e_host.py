"""Where the time of Retriever.retrieve goes at C2 with host inputs (cProfile of the host side + wall time per call)."""
import cProfile
import pstats
import sys
import time

import torch

sys.path.insert(0, ".")
from rag_docvqa_b200 import synth  # noqa: E402
from rag_docvqa_b200.retriever import Retriever  # noqa: E402

dev = torch.device("cuda:0")
w = synth.WORKLOADS["C2"]
hb = synth.make_text_batch("C2", with_lists=True, share_image_pool=24)
emb = [e.cpu().pin_memory() for e in hb["text_embeddings"]]
q = hb["question_embeddings"].cpu().pin_memory()
lists = (hb["words_text_chunks"], hb["words_box_chunks"], hb["layout_labels_chunks"], hb["images"], hb["page_indices"])
cached = len(sys.argv) > 1 and sys.argv[1] == "cached"
retr = Retriever({"compute_stats": False, "compute_stats_examples": False, "n_stats_examples": 0, "chunk_num": w.k, "device": str(dev),
                  "retrieval_lazy_patches": True, "retrieval_pause_gc": True, "retrieval_embedding_cache_mb": 8192 if cached else 0})
for _ in range(5):
    retr.retrieve(emb, q, *lists)
n = 40
t0 = time.perf_counter()
for _ in range(n):
    retr.retrieve(emb, q, *lists)
dt = (time.perf_counter() - t0) / n
print("retrieve: %.3f ms per call = %.0f questions/s" % (dt * 1e3, w.docs / dt))
pr = cProfile.Profile()
pr.enable()
for _ in range(n):
    retr.retrieve(emb, q, *lists)
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(18)
