"""Which part of the input-embedding kernel costs the time: spatial only (no token rows), all tokens one box (no coordinate
rows after the first of a chunk), both."""
import sys

import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
from rag_docvqa_b200.vt5_embed import SpatialEmbeddings, VT5InputEmbeddings  # noqa: E402

dev = torch.device("cuda:0")
B, L, D, V = 64, 512, 768, 32128
g = torch.Generator().manual_seed(1)
sp = SpatialEmbeddings(torch.randn(1024, D, generator=g), torch.randn(1024, D, generator=g), torch.ones(D), torch.zeros(D), 1e-12,
                       torch.randn(D, D, generator=g) / 28, torch.zeros(D), device=dev)
emb = VT5InputEmbeddings(sp, torch.randn(V, D, generator=g))
ids = torch.randint(0, V, (B, L), generator=g).to(dev)
ids_same = torch.zeros((B, L), dtype=torch.int64, device=dev)
box_rand = torch.randint(0, 1001, (B, L, 4), generator=g).to(dev)
box_same = torch.zeros((B, L, 4), dtype=torch.int64, device=dev)


def t(fn):
    for _ in range(3):
        fn()
    return bench.timed_loop(lambda i: fn(), 20, torch.cuda.synchronize) / 20 * 1e3


print("random ids, random boxes   %.1f us" % t(lambda: emb(ids, box_rand)))
print("random ids, one box        %.1f us" % t(lambda: emb(ids, box_same)))
print("one id,     random boxes   %.1f us" % t(lambda: emb(ids_same, box_rand)))
print("one id,     one box        %.1f us" % t(lambda: emb(ids_same, box_same)))
print("spatial only, random boxes %.1f us" % t(lambda: sp(box_rand)))
print("spatial only, one box      %.1f us" % t(lambda: sp(box_same)))
out = torch.empty((B, L, D), device=dev)
print("torch fill of the output   %.1f us" % t(lambda: out.fill_(1.0)))
