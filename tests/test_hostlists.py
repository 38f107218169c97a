"""CPU checks of the host-side list view (csrc/hostlists.c): same output as the Python walk in
retriever.Retriever._hit_lists and as the oracle's gather_hits (reference src/_modules.py:2014-2153)."""
import numpy as np
import pytest

from oracle import ref_restated as R
from rag_docvqa_b200 import synth


def _retriever(lazy=True, s=0, reorder=False):
    from rag_docvqa_b200 import retriever as RT

    class Host(RT.Retriever):          # list building needs no device: skip the CUDA-only constructor
        def __init__(self):
            self.include_surroundings, self.reorder_chunks, self.lazy_patches = s, reorder, lazy
    return RT, Host()


def _batch():
    b = synth.make_text_batch("C2", with_lists=True, docs=12, seed=91, share_image_pool=6)
    return (b["words_text_chunks"], b["words_box_chunks"], b["layout_labels_chunks"], b["images"], b["page_indices"]), b["sizes"]


def _hits(sizes, k, seed):
    rng = np.random.RandomState(seed)
    return [[int(i) for i in rng.permutation(n)[:min(k, n)]] for n in sizes]


def test_module_is_built_and_used():
    from rag_docvqa_b200 import build
    build.build_hostlists()
    RT, _ = _retriever()
    assert RT._hostlists is not None and hasattr(RT._hostlists, "gather_s0")


@pytest.mark.parametrize("k", [1, 5, 20])
def test_c_walk_equals_python_walk_and_oracle(k):
    RT, r = _retriever()
    args, sizes = _batch()
    hits = _hits(sizes, k, 3 + k)
    out_c = r._hit_lists(hits, *args)
    saved, RT._hostlists = RT._hostlists, None
    try:
        out_py = r._hit_lists(hits, *args)
    finally:
        RT._hostlists = saved
    ref = R.gather_hits(hits, *args, include_surroundings=0, reorder_chunks=False, crop=False)
    for i in (0, 1, 2, 3, 4, 5, 7):
        assert list(out_c[i]) == out_py[i] == ref[i], "output %d differs" % i
    rects_c = [[list(im._rect) for im in doc] for doc in out_c[6]]
    assert rects_c == [[list(im._rect) for im in doc] for doc in out_py[6]] == ref[6]
    # identity, like the reference: the emitted box objects are the caller's objects, the lists are fresh
    b0 = next(b for b, h in enumerate(hits) if h)
    assert out_c[4][b0][0][0] is args[1][b0][hits[b0][0]][0]
    assert out_c[4][b0][0] is not args[1][b0][hits[b0][0]]


def test_python_semantics_on_odd_inputs():
    from PIL import Image
    RT, r = _retriever(lazy=False)
    page = Image.new("RGB", (200, 100))
    words = [[["a", "b"], [], ["c"], ["d", "e", "f"]]]
    boxes = [[[(0.5, 0.25, 0.75, 0.5), [0.1, 0.3, 0.2, 0.35]],                 # tuple + list boxes
              [],                                                               # empty chunk -> [0, 0, 1, 1], whole page
              [[1, 0, 0, 1]],                                                   # ints, x0 > x1 -> order fix
              [[0.2, 0.2, 0.4, 0.4], [0.2, 0.1, 0.4, 0.4], [0.2, 0.2, 0.4, 0.4]]]]   # ties: first extreme kept
    labels, pages, images = [[3, 1, 2, 0]], [[0, 0, 0, 0]], [[page]]
    hits = [[3, 1, 2, 0]]
    out = r._hit_lists(hits, words, boxes, labels, images, pages)
    ref = R.gather_hits(hits, words, boxes, labels, images, pages, include_surroundings=0, reorder_chunks=False)
    for i in (0, 1, 2, 3, 4, 5, 7):
        assert list(out[i]) == ref[i], "output %d differs" % i
    assert [im.size for im in out[6][0]] == [im.size for im in ref[6][0]]
    assert out[1][0][1] == [0, 0, 1, 1] and out[6][0][1].size == (200, 100)
    assert out[1][0][0][0] is boxes[0][3][0][0]                                 # min() returns the FIRST minimal element
    with pytest.raises(TypeError):
        r._hit_lists([[0]], [[["a", 5]]], [[[[0, 0, 1, 1], [0, 0, 1, 1]]]], [[1]], images, [[0]])   # " ".join of a non-str
    with pytest.raises(IndexError):
        r._hit_lists([[7]], words, boxes, labels, images, pages)


def test_numpy_hits_and_labels():
    RT, r = _retriever()
    args, sizes = _batch()
    hits = _hits(sizes, 4, 11)
    words, boxes, labels, images, pages = args
    out_a = r._hit_lists(hits, *args)
    out_b = r._hit_lists([np.asarray(h, dtype=np.int64) for h in hits], words, boxes,
                         [np.asarray(x) for x in labels], images, [np.asarray(x) for x in pages])
    for i in (0, 1, 3, 4):
        assert list(out_a[i]) == list(out_b[i])
    assert [[int(x) for x in d] for d in out_b[2]] == out_a[2]
    assert [[int(x) for x in d] for d in out_b[7]] == out_a[7]


def test_lazy_crop_behaves_like_an_eager_crop():
    """`retrieval_lazy_patches`: the deferred patch must be indistinguishable from page.crop(rect)
    (reference src/_modules.py:2119) through PIL's public operations."""
    from PIL import Image
    from rag_docvqa_b200.retriever import lazy_crop
    page = Image.fromarray(np.random.RandomState(0).randint(0, 255, (100, 200, 3)).astype(np.uint8), "RGB")
    rect = (10, 20, 110, 90)
    ref = page.crop(rect)
    assert lazy_crop(page, rect).size == ref.size and lazy_crop(page, rect).mode == ref.mode
    assert lazy_crop(page, rect).tobytes() == ref.tobytes()
    assert np.array_equal(np.asarray(lazy_crop(page, rect)), np.asarray(ref))
    assert lazy_crop(page, rect).resize((50, 40)).tobytes() == ref.resize((50, 40)).tobytes()
    assert lazy_crop(page, rect).convert("L").tobytes() == ref.convert("L").tobytes()
    assert lazy_crop(page, rect).copy().tobytes() == ref.tobytes()
    assert lazy_crop(page, (0, 0, 0, 0)).size == (0, 0)


def test_pix2struct_plan_matches_the_reference_arithmetic():
    """Host half of rdv_pix2struct_patches: patch grids, output offsets and row-id offsets of
    extract_multi_image_flattened_patches (src/custom_pix2struct_processor.py:52-57, 97-132), checked against the
    oracle's run of the same images."""
    from oracle import ref_restated as R
    from rag_docvqa_b200.pagestore import plan_pix2struct
    rng = np.random.RandomState(12)
    for max_total in (2048, 300, 64):
        sizes = [(int(rng.randint(8, 400)), int(rng.randint(8, 900))) for _ in range(rng.randint(1, 7))]     # (h, w)
        crops = [[(i, 0, 0, w, h) for i, (h, w) in enumerate(sizes)]]
        images, doc_total, temp = plan_pix2struct(crops, np.array([0, len(sizes)], dtype=np.int32), max_total, 16)
        per, row_offset, out_start = max_total // len(sizes), 0, 0
        for rec, (h, w) in zip(images, sizes):
            flat, nxt = R.pix2struct_patches_single(np.zeros((h, w, 3), np.float32), per, 16, 16, row_offset)
            assert rec["kept"] == flat.shape[0] and rec["out_start"] == out_start and rec["row_offset"] == row_offset
            assert rec["rows"] == nxt - row_offset and rec["rows"] * rec["cols"] >= rec["kept"]
            out_start += flat.shape[0]
            row_offset = nxt
        assert doc_total.tolist() == [out_start] and out_start <= max_total


def test_docstore_arrays_c_walk_equals_python_walk():
    """DocStore.arrays_from_lists: the per-word walk in C (flatten_docs) and in Python give identical CSR arrays."""
    from rag_docvqa_b200 import docstore
    b = synth.make_text_batch("C2", with_lists=True, docs=6, seed=19, share_image_pool=4)
    words = b["words_text_chunks"]
    words[1][3] = []                                             # an empty chunk
    b["words_box_chunks"][1][3] = []
    table = synth.make_tokens_for_words(words, seed=3)
    args = (words, b["words_box_chunks"], b["layout_labels_chunks"], b["page_indices"], lambda w: table.get(w, [2]))
    assert docstore._hostlists is not None
    a_c, B_c = docstore.DocStore.arrays_from_lists(*args, images=b["images"])
    saved, docstore._hostlists = docstore._hostlists, None
    try:
        a_py, B_py = docstore.DocStore.arrays_from_lists(*args, images=b["images"])
    finally:
        docstore._hostlists = saved
    assert B_c == B_py == 6 and sorted(a_c) == sorted(a_py)
    for k in a_c:
        assert np.array_equal(a_c[k], a_py[k]), k
    bad_boxes = [list(d) for d in b["words_box_chunks"]]
    bad_boxes[0] = list(bad_boxes[0])
    bad_boxes[0][0] = bad_boxes[0][0][:-1]                      # one box short
    with pytest.raises(ValueError):
        docstore.DocStore.arrays_from_lists(words, bad_boxes, b["layout_labels_chunks"], b["page_indices"], lambda w: [2])
