"""Visual pack on the C2 batch (ncu target): 64 documents x 5 retrieved chunks -> 224 x 224 inputs."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rag_docvqa_b200 import synth
from rag_docvqa_b200.docstore import DocStore
from rag_docvqa_b200.pagestore import PageStore
from rag_docvqa_b200.retriever import Retriever
dev = torch.device("cuda:0")
b = synth.make_text_batch("C2", with_lists=True, share_image_pool=24)
table = synth.make_tokens_for_words(b["words_text_chunks"], seed=3)
store = DocStore.from_lists(b["words_text_chunks"], b["words_box_chunks"], b["layout_labels_chunks"], b["page_indices"],
                            lambda w: table.get(w, [2]), dev, images=b["images"])
pstore = PageStore.from_images(b["images"], dev)
r = Retriever({"compute_stats": False, "compute_stats_examples": False, "n_stats_examples": 0, "chunk_num": 5, "device": "cuda:0"})
pk, rs = r.retrieve_packed([e.to(dev) for e in b["text_embeddings"]], b["question_embeddings"].to(dev), store, [[5, 6]] * 64)
plan = pstore.prepare_pack(pk.hit_page, pk.hit_rect, rs.topk_cnt)
for _ in range(3):
    plan.launch()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    plan.launch()
e1.record()
torch.cuda.synchronize()
print("visual pack: %.3f ms per batch" % (e0.elapsed_time(e1) / 5))
