"""TEST INFRASTRUCTURE ONLY -- runs the UNMODIFIED wrapper RAGVT5.online_retrieve (src/RAGVT5.py:153-316) of the reference
on a synthetic batch, with the reference's own Chunker and Retriever and a stand-in embedder (the BiEncoder is a model and out
of scope), records the arguments the wrapper hands to Retriever.retrieve (:244-252) and freezes them with the wrapper's
outputs -> tests/golden/wrapper_online_retrieve.npz / .json.

    PYTHONDONTWRITEBYTECODE=1 python -m oracle.make_golden_wrapper        (build container: /root/reference must exist)
"""
import importlib
import json
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.ref_import import import_reference  # noqa: E402
from rag_docvqa_b200 import synth  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")

CONFIG = {"compute_stats": True, "compute_stats_examples": False, "n_stats_examples": 0, "chunk_num": 4, "chunk_size": 30,
          "chunk_size_tol": 0.2, "overlap": 5, "include_surroundings": 0, "reorder_chunks": False, "page_retrieval": "concat",
          "layout_model": "", "layout_model_weights": "", "device": "cuda:0"}


def import_ragvt5():
    modules, utils, model_utils = import_reference()
    if "src.QwenVLInstruct" not in sys.modules:        # src/RAGVT5.py:20: dependencies absent here, never touched by retrieval
        stub = types.ModuleType("src.QwenVLInstruct")
        stub.QwenVLForConditionalGeneration = None
        sys.modules["src.QwenVLInstruct"] = stub
    return modules, importlib.import_module("src.RAGVT5")


def make_batch(seed=5, docs=5):
    """What the dataset hands the wrapper (src/RAGVT5.py:170-175): questions, per-page words / boxes, page images."""
    from PIL import Image
    rng = np.random.RandomState(seed)
    batch = {"questions": [], "words": [], "boxes": [], "images": [], "question_id": list(range(docs))}
    for b in range(docs):
        n_pages = [3, 1, 2, 1, 4][b % 5]
        words_b, boxes_b, images_b = [], [], []
        for p in range(n_pages):
            n = 0 if (b == 3) else int(rng.randint(20, 90))                  # document 3: no OCR words at all
            words_b.append(["w%d" % i for i in rng.randint(0, 400, size=n)])
            x0, y0 = rng.uniform(0, 0.9, size=n), rng.uniform(0, 0.95, size=n)
            boxes_b.append(np.stack([x0, y0, x0 + rng.uniform(0.005, 0.1, size=n), y0 + rng.uniform(0.005, 0.05, size=n)], axis=1).tolist())
            images_b.append(Image.fromarray(rng.randint(0, 256, (110 + 3 * p, 85 + 2 * b, 3)).astype(np.uint8), "RGB"))
        batch["questions"].append("what is item %d about ?" % b)
        batch["words"].append(words_b); batch["boxes"].append(boxes_b); batch["images"].append(images_b)
    return batch


class Embedder:
    """Stand-in for BiEncoder (src/_modules.py:1415-1477): batch_forward -> one (n_b, d) tensor per document, forward -> (B, d)."""

    def __init__(self, dim=96):
        self.h = synth.HashEmbedder(dim)

    def batch_forward(self, text_chunks):
        return [self.h.forward(list(doc)) for doc in text_chunks]

    def forward(self, texts):
        return self.h.forward(list(texts))


def stand_in_self(ragvt5, modules, config):
    """The attributes online_retrieve reads (src/RAGVT5.py:170-316), built the way RAGVT5.__init__ builds them (:96-105):
    the Chunker and the Retriever come from the names bound INSIDE src.RAGVT5, which compat.install() rebinds."""
    return types.SimpleNamespace(
        layout_model=None, use_precomputed_layouts=False, chunker=ragvt5.Chunker(config), use_layout_labels="Default",
        layout_map=modules.get_layout_model_map(config), embedder=Embedder(), retriever=ragvt5.Retriever(config),
        reranker=None, page_retrieval=config["page_retrieval"], train_mode=False, train_layout=False, train_embedder=False)


def jsonable(x):
    if isinstance(x, (list, tuple)):
        return [jsonable(v) for v in x]
    if isinstance(x, (np.integer,)):
        return int(x)
    if isinstance(x, (np.floating,)):
        return float(x)
    return x


def main():
    modules, ragvt5 = import_ragvt5()
    batch = make_batch()
    me = stand_in_self(ragvt5, modules, CONFIG)
    seen = {}
    inner = me.retriever.retrieve

    def spy(*args):
        seen["args"] = args
        return inner(*args)
    me.retriever.retrieve = spy
    out = ragvt5.RAGVT5.online_retrieve(me, batch)                    # the reference wrapper, unmodified
    emb, q, words_chunks, boxes_chunks, labels_chunks, images, page_indices = seen["args"]
    arrays = {"q": q.numpy(), "docs": np.int64(len(emb))}
    for b, e in enumerate(emb):
        arrays["emb_%d" % b] = e.numpy()
        arrays["sims_%d" % b] = out[9][b].numpy()
        for p, im in enumerate(images[b]):
            arrays["page_%d_%d" % (b, p)] = np.asarray(im)
        for j, patch in enumerate(out[3][b]):
            arrays["patch_%d_%d" % (b, j)] = np.asarray(patch)
    np.savez_compressed(os.path.join(GOLDEN, "wrapper_online_retrieve.npz"), **arrays)
    frozen = {
        "config": CONFIG, "n_pages": [len(p) for p in images],
        "retrieve_args": {"words_text_chunks": words_chunks, "words_box_chunks": jsonable(boxes_chunks),
                          "layout_labels_chunks": jsonable(labels_chunks), "page_indices": jsonable(page_indices)},
        "outputs": {"top_k_text": out[0], "top_k_boxes": jsonable(out[1]), "top_k_layout_labels": jsonable(out[2]),
                    "n_patches": [len(p) for p in out[3]], "top_k_page_indices": jsonable(out[4]), "top_k_words_text": out[5],
                    "top_k_words_boxes": jsonable(out[6]), "top_k_words_layout_labels": jsonable(out[7]),
                    "words_layout_labels_pages": jsonable(out[8])},
        "retriever_stats": jsonable(dict(me.retriever.stats.get("layout_labels_topk_dist", {}))),
    }
    with open(os.path.join(GOLDEN, "wrapper_online_retrieve.json"), "w") as f:
        json.dump(frozen, f)
    mpath = os.path.join(GOLDEN, "MANIFEST.json")
    manifest = json.load(open(mpath))
    what = ("RAGVT5.online_retrieve of the reference (src/RAGVT5.py:153-316), unmodified, with its own Chunker / Retriever and a "
            "stand-in embedder: the arguments it hands to Retriever.retrieve and its outputs")
    manifest["files"]["wrapper_online_retrieve.npz"] = what
    manifest["files"]["wrapper_online_retrieve.json"] = what
    json.dump(manifest, open(mpath, "w"), indent=1)
    print("wrote wrapper_online_retrieve.{npz,json}:", os.path.getsize(os.path.join(GOLDEN, "wrapper_online_retrieve.npz")),
          os.path.getsize(os.path.join(GOLDEN, "wrapper_online_retrieve.json")), "bytes; hits per document:", [len(t) for t in out[0]])


if __name__ == "__main__":
    main()
