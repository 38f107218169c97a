"""Pix2Struct patch assembly of 8 documents x 5 strips of 850 x 220 (ncu target)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from PIL import Image
from rag_docvqa_b200.pagestore import PageStore
dev = torch.device("cuda:0")
rng = np.random.RandomState(7)
B, k = 8, 5
pages = [[Image.fromarray(rng.randint(0, 256, (220, 850, 3)).astype(np.uint8), "RGB") for _ in range(k)] for _ in range(B)]
store = PageStore.from_images(pages, dev)
crops = [[(g, 0, 0, 850, 220) for g in range(k)] for _ in range(B)]
for _ in range(3):
    store.pack_pix2struct(crops)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    store.pack_pix2struct(crops)
e1.record(); torch.cuda.synchronize()
print("pix2struct patches: %.3f ms per batch of %d documents" % (e0.elapsed_time(e1) / 5, B))
