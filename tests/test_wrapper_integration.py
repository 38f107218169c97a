"""The drop-in behind the reference's own, UNMODIFIED wrapper: RAGVT5.online_retrieve (src/RAGVT5.py:153-316).

  * CPU, build container only (needs /root/reference): the wrapper is run twice on the same batch -- with the reference's
    Retriever, and after rag_docvqa_b200.compat.install() with the drop-in, whose device round trip is emulated on the host
    (tests/_emulated_device.py: real layout / pack / list building, the oracle in place of the kernel) -- and every output of
    the wrapper must agree.  This exercises exactly what "swap in unchanged" means: the constructor call at :105, the
    arguments of :244-252, the 9-tuple unpacked at :233-243, `retriever.stats` read at :294.
  * `-m gpu` (no reference on the GPU box): the arguments the unmodified wrapper handed to Retriever.retrieve and its
    outputs were frozen by oracle/make_golden_wrapper.py; the drop-in on the device must reproduce the outputs from them.
"""
import json
import os

import numpy as np
import pytest
import torch

from oracle.ref_import import reference_available


def load_frozen(golden_dir):
    from PIL import Image
    z = np.load(os.path.join(golden_dir, "wrapper_online_retrieve.npz"))
    with open(os.path.join(golden_dir, "wrapper_online_retrieve.json")) as f:
        fz = json.load(f)
    docs = int(z["docs"])
    emb = [torch.from_numpy(z["emb_%d" % b]) for b in range(docs)]
    q = torch.from_numpy(z["q"])
    images = [[Image.fromarray(z["page_%d_%d" % (b, p)], "RGB") for p in range(fz["n_pages"][b])] for b in range(docs)]
    return z, fz, emb, q, images


def check_outputs(out9, z, fz, device_sims=False):
    o = fz["outputs"]
    assert out9[0] == o["top_k_text"]
    assert out9[1] == o["top_k_boxes"]
    assert out9[2] == o["top_k_layout_labels"]
    assert out9[3] == o["top_k_words_text"]
    assert out9[4] == o["top_k_words_boxes"]
    assert out9[5] == o["top_k_words_layout_labels"]
    assert out9[7] == o["top_k_page_indices"]
    assert [len(p) for p in out9[6]] == o["n_patches"]
    for b, patches in enumerate(out9[6]):
        for j, patch in enumerate(patches):
            assert np.array_equal(np.asarray(patch), z["patch_%d_%d" % (b, j)])
    for b, s in enumerate(out9[8]):
        ref = z["sims_%d" % b]
        got = s.cpu().numpy()
        assert got.shape == ref.shape
        if device_sims:
            np.testing.assert_allclose(got, ref, rtol=1e-5, atol=1e-6)
        else:
            assert np.array_equal(got, ref)


@pytest.mark.skipif(not reference_available(), reason="needs the reference tree (build container)")
def test_unmodified_wrapper_with_and_without_the_drop_in(golden_dir, monkeypatch):
    import _emulated_device
    from oracle import make_golden_wrapper as W
    from rag_docvqa_b200 import compat
    from rag_docvqa_b200.retriever import Retriever
    modules, ragvt5 = W.import_ragvt5()
    batch = W.make_batch()
    # 1. the reference as it is
    me = W.stand_in_self(ragvt5, modules, W.CONFIG)
    assert type(me.retriever).__module__ == "src._modules"
    ref = ragvt5.RAGVT5.online_retrieve(me, batch)
    # 2. the same wrapper code after compat.install(): the names bound inside src.RAGVT5 now build the drop-ins
    _emulated_device.install(monkeypatch)
    patched = compat.install()
    try:
        assert ("src.RAGVT5", "Retriever") in patched
        me2 = W.stand_in_self(ragvt5, modules, W.CONFIG)
        assert isinstance(me2.retriever, Retriever) and type(me2.chunker).__module__ == "rag_docvqa_b200.chunker"
        new = ragvt5.RAGVT5.online_retrieve(me2, batch)
    finally:
        compat.uninstall()
    assert ragvt5.Retriever is modules.Retriever                     # uninstall restored the reference
    for i in (0, 1, 2, 4, 5, 6, 7, 8):                               # texts, boxes, labels, pages, words, word boxes, word labels
        assert new[i] == ref[i], i
    assert [len(p) for p in new[3]] == [len(p) for p in ref[3]]
    for pa, pb in zip(new[3], ref[3]):
        for a, b in zip(pa, pb):
            assert np.array_equal(np.asarray(a), np.asarray(b))      # the crops, pixel for pixel
    for a, b in zip(new[9], ref[9]):
        assert torch.equal(a.cpu(), b.cpu())                          # every similarity
    assert new[11]["stats"]["layout_labels_topk_dist"] == ref[11]["stats"]["layout_labels_topk_dist"] \
        if "layout_labels_topk_dist" in ref[11]["stats"] else True
    # ... and both equal what oracle/make_golden_wrapper.py froze for the GPU test below
    z, fz, _, _, _ = load_frozen(golden_dir)
    check_outputs((ref[0], ref[1], ref[2], ref[5], ref[6], ref[7], ref[3], ref[4], ref[9]), z, fz)


@pytest.mark.gpu
@pytest.mark.parametrize("where", ["device", "host"])
def test_drop_in_reproduces_the_wrapper_outputs(golden_dir, where):
    """What the unmodified RAGVT5.online_retrieve handed to Retriever.retrieve, and what it returned (frozen in the build
    container), against the drop-in on the device."""
    from rag_docvqa_b200.retriever import Retriever
    z, fz, emb, q, images = load_frozen(golden_dir)
    a = fz["retrieve_args"]
    retr = Retriever({**fz["config"], "device": "cuda:0"})
    if where == "device":
        emb, q = [e.to("cuda:0") for e in emb], q.to("cuda:0")
    out = retr.retrieve(emb, q, a["words_text_chunks"], a["words_box_chunks"], a["layout_labels_chunks"], images, a["page_indices"])
    check_outputs(out, z, fz, device_sims=True)
    assert all(s.is_cuda == (where == "device") for s in out[8])      # similarities live where the embeddings live


@pytest.mark.skipif(not reference_available(), reason="needs the reference tree (build container)")
def test_unmodified_pix2struct_wrapper_with_and_without_the_drop_in(monkeypatch):
    """RAGPix2Struct.online_retrieve (src/RAGPix2Struct.py:104-181), unmodified: the reference's ImageChunker cuts the pages
    into strips, a stand-in encoder embeds them, and the retriever -- the reference's VisualRetriever, then the drop-in after
    compat.install() -- picks the strips.  The drop-in's two device calls are replaced by the oracle (no GPU here); the
    decode, the neighbourhoods, the merge of overlapping rectangles and the crops are its own code.  Crops and page ids are
    compared as multisets (the reference's order is Python-set iteration order, src/_modules.py:2428,2445)."""
    import importlib
    import sys
    import types
    import zlib

    from PIL import Image

    from oracle import ref_restated as R
    from oracle.ref_import import import_reference
    from rag_docvqa_b200 import compat
    from rag_docvqa_b200 import functional as F
    from rag_docvqa_b200 import retriever as RM
    modules, _, _ = import_reference()
    # names src/RAGPix2Struct.py:13-16 imports and online_retrieve never touches (generator-side processors); render_text only
    # feeds the stand-in encoder
    stub = sys.modules.get("src.custom_pix2struct_processor") or types.ModuleType("src.custom_pix2struct_processor")
    for name in ("CustomPix2StructProcessor", "CustomPix2StructImageProcessor"):
        if not hasattr(stub, name):
            setattr(stub, name, None)
    sys.modules["src.custom_pix2struct_processor"] = stub
    rag = importlib.import_module("src.RAGPix2Struct")
    # render_text (transformers) fetches a font from the hub: there is no network here, and the question image only feeds the
    # stand-in encoder
    fake_render = lambda text, **kw: Image.new("RGB", (64, 16), (len(text) % 255, zlib.crc32(text.encode()) % 255, 0))

    class Encoder:
        """Stand-in for ImageEncoder (src/_modules.py:1627-1666): (n_strips, L, d) per document, zeros(1, L, d) for none."""
        L, d = 24, 32

        def one(self, im):
            seed = zlib.crc32(np.asarray(im).tobytes()) & 0x7FFFFFFF
            return torch.randn(self.L, self.d, generator=torch.Generator().manual_seed(seed))

        def batch_forward(self, docs):
            return [torch.stack([self.one(im) for im in doc]) if len(doc) else torch.zeros(1, self.L, self.d) for doc in docs]

        def forward(self, images):
            return torch.stack([self.one(im) for im in images])

    rng = np.random.RandomState(3)
    images = [[Image.fromarray(rng.randint(0, 256, (300 + 40 * p, 200, 3)).astype(np.uint8), "RGB") for p in range(n)] for n in (2, 1, 3)]
    batch = {"questions": ["what is item %d ?" % b for b in range(3)], "images": images, "question_id": [0, 1, 2]}
    out = {}
    for surroundings in (0, 1):
        config = {"chunk_num": 3, "include_surroundings": surroundings, "chunk_mode": "horizontal", "patch_size": 96, "overlap": 0,
                  "layout_model": "", "device": "cuda:0"}

        def me():
            return types.SimpleNamespace(layout_model=None, use_precomputed_layouts=False, chunker=rag.ImageChunker(config),
                                         embedder=Encoder(), retriever=rag.VisualRetriever(config))
        monkeypatch.setattr(rag, "render_text", fake_render)
        ref_self = me()
        assert type(ref_self.retriever).__module__ == "src._modules"
        ref = rag.RAGPix2Struct.online_retrieve(ref_self, batch)
        # the drop-in's device calls, answered by the oracle
        monkeypatch.setattr(RM.VisualRetriever, "_get_similarities",
                            lambda self, patches, q: [R.late_interaction(q[i].unsqueeze(0), patches[i]) for i in range(len(patches))])
        monkeypatch.setattr(RM, "_to_device", lambda t, dev: t)

        def topk_segments(scores, k):
            B = len(scores)
            idx = torch.full((B, k), -1, dtype=torch.int32)
            cnt = torch.zeros(B, dtype=torch.int32)
            for b, s in enumerate(scores):
                hits = R.topk_lowest_index(s, k)
                idx[b, :len(hits)] = torch.from_numpy(np.asarray(hits, dtype=np.int32))
                cnt[b] = len(hits)
            return idx, None, cnt
        monkeypatch.setattr(F, "topk_segments", topk_segments)
        patched = compat.install()
        try:
            assert ("src.RAGPix2Struct", "VisualRetriever") in patched
            new_self = me()
            assert isinstance(new_self.retriever, RM.VisualRetriever)
            new = rag.RAGPix2Struct.online_retrieve(new_self, batch)
        finally:
            compat.uninstall()
            monkeypatch.undo()
        for b in range(3):
            assert sorted(new[1][b]) == sorted(ref[1][b])                                       # page ids
            key = lambda im: (im.size, np.asarray(im).tobytes())
            assert sorted(map(key, new[0][b])) == sorted(map(key, ref[0][b]))                   # the crops, pixel for pixel
        out[surroundings] = [len(c) for c in new[0]]
    assert all(n > 0 for n in out[0])
