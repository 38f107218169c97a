"""CPU checks of the host half of rdv_retrieve_small_f32 (include/rdv.h): rdv_small_batch_layout and
rdv_small_batch_pack are pure host code -- the packed upload blob must describe the batch exactly as the general path's
rdv_build_doc_table does, and scoring the BLOB's rows with the oracle must equal scoring the documents."""
import ctypes

import numpy as np
import pytest
import torch

from oracle import ref_restated as R
from rag_docvqa_b200 import _lib, synth
from rag_docvqa_b200.functional import CTA_DTYPE, TILE_DTYPE

CASES = [([30], 384, 5), ([45, 0, 3], 384, 5), ([60, 0, 3, 150, 1, 90, 31], 768, 10), ([0, 0], 8, 1), ([700], 128, 40)]


def pack(sizes, d, k, seed=3):
    emb, q = synth.make_embeddings(sizes, d, seed)
    B = len(sizes)
    rows = (ctypes.c_int64 * B)(*sizes)
    lay = _lib.SmallLayoutStruct()
    _lib.check(_lib.lib.rdv_small_batch_layout(rows, B, d, k, ctypes.addressof(lay)))
    h_blob = np.zeros(lay.in_bytes + 16, dtype=np.uint8)
    base = h_blob.ctypes.data + (-h_blob.ctypes.data) % 16
    h_docs = (ctypes.c_void_p * B)(*[e.data_ptr() if n else None for e, n in zip(emb, sizes)])
    fake_dev = 0x7F0000000000                       # the tile descriptors point into the DEVICE twin of the blob
    _lib.check(_lib.lib.rdv_small_batch_pack(h_docs, rows, B, d, q.data_ptr(), ctypes.addressof(lay), base, fake_dev))
    off = base - h_blob.ctypes.data
    return emb, q, lay, h_blob[off:off + lay.in_bytes], fake_dev


def check_cluster_table(ctas, sizes, row_off, srcs, cs, slice_rows=32):
    """rdv_build_cluster_table: every document owns min(max(ceil(n / slice_rows), 1), cluster size) CONSECUTIVE CTAs of ONE
    cluster, parts numbered from 0; everything else is padding (nparts == 0)."""
    seen = {}
    for i, c in enumerate(ctas):
        if int(c["nparts"]) == 0:
            continue
        b = int(c["doc"])
        seen.setdefault(b, []).append(i)
        assert int(c["doc_rows"]) == sizes[b] and int(c["sims_off"]) == row_off[b] and int(c["src"]) == srcs[b]
    assert sorted(seen) == [b for b in range(len(sizes))]
    for b, where in seen.items():
        want = min(max(-(-sizes[b] // slice_rows), 1), cs)
        assert len(where) == want and where == list(range(where[0], where[0] + want))          # consecutive
        assert where[0] // cs == where[-1] // cs                                                # one cluster
        assert [int(ctas[i]["part"]) for i in where] == list(range(want))
        assert all(int(ctas[i]["nparts"]) == want for i in where)


@pytest.mark.parametrize("slice_rows", [32, 40, 64])
@pytest.mark.parametrize("cs", [8, 16])
def test_cluster_table_packing(cs, slice_rows):
    rng = np.random.RandomState(4)
    assert _lib.lib.rdv_cluster_max_rows(cs) == 4 * 32 * cs
    for B in (1, 2, 7, 64, 300):
        sizes = rng.randint(0, _lib.lib.rdv_cluster_max_rows(cs) + 1, size=B).astype(np.int64)
        sizes[0] = 0
        if B > 2:
            sizes[1] = 600
            sizes[2] = 32 * cs
        n = _lib.lib.rdv_cluster_table_size(sizes.ctypes.data, B, cs, slice_rows)
        assert n > 0 and n % cs == 0
        slots = int(np.minimum(np.maximum(-(-sizes // slice_rows), 1), cs).sum())
        assert n >= slots and n - slots < max(cs, 0.25 * slots + cs)          # best-fit decreasing wastes little
        ctas = np.zeros(n, dtype=CTA_DTYPE)
        srcs = [(0x10000 + 16 * b) if sizes[b] else 0 for b in range(B)]
        ptrs = np.asarray(srcs, dtype=np.uint64)
        _lib.check(_lib.lib.rdv_build_cluster_table(ptrs.ctypes.data, sizes.ctypes.data, B, 64, cs, slice_rows, ctas.ctypes.data, n))
        row_off = np.concatenate([[0], np.cumsum(sizes)])
        check_cluster_table(ctas, sizes.tolist(), row_off, srcs, cs, slice_rows)
    big = np.asarray([_lib.lib.rdv_cluster_max_rows(cs) + 1], dtype=np.int64)
    ptrs = np.asarray([0x10000], dtype=np.uint64)
    ctas = np.zeros(cs, dtype=CTA_DTYPE)
    assert _lib.lib.rdv_build_cluster_table(ptrs.ctypes.data, big.ctypes.data, 1, 64, cs, slice_rows, ctas.ctypes.data, cs) == _lib.E_LIMIT
    assert _lib.lib.rdv_cluster_table_size(big.ctypes.data, 1, 12, 32) == -1                  # only clusters of 8 and 16 exist
    assert _lib.lib.rdv_cluster_table_size(big.ctypes.data, 1, cs, 36) == -1                  # slices: multiples of 8 in [32, 64]


@pytest.mark.parametrize("sizes,d,k", CASES)
def test_layout_and_pack(sizes, d, k):
    emb, q, lay, blob, dev = pack(sizes, d, k)
    B, total = len(sizes), sum(sizes)
    assert lay.total_rows == total and lay.max_rows == max(sizes)
    cs, slice_rows = 8, 32                                                         # rdv_cluster_plan without a device
    cluster = (max(sizes) <= _lib.lib.rdv_cluster_max_rows(8) and k <= _lib.lib.rdv_cluster_max_k()
               and (total + B) * d * 4 <= (1 << 20))                                # rdv_retrieve_plan
    assert lay.algo == (_lib.SMALL_CLUSTER if cluster else _lib.SCORE_LDG) and 1 <= lay.tile_rows <= 1024
    assert lay.n_tiles == sum(-(-n // lay.tile_rows) for n in sizes)
    # every part 16-byte aligned, parts in order and not overlapping
    assert lay.o_tiles % 32 == 0 and lay.o_q % 16 == 0 and lay.o_emb % 16 == 0
    assert 8 * (B + 1) <= lay.o_tiles and lay.o_tiles + 32 * lay.n_tiles <= lay.o_ctas and lay.o_ctas + 32 * lay.n_ctas == lay.o_q
    assert (lay.n_ctas > 0) == cluster and lay.n_ctas % cs == 0
    assert (lay.cluster, lay.slice_rows) == ((cs, slice_rows) if cluster else (0, 0))
    assert lay.o_q + B * d * 4 == lay.o_emb and lay.o_emb + total * d * 4 == lay.in_bytes
    assert lay.o_idx == total * 4 and lay.o_cnt == lay.o_idx + B * k * 4 and lay.read_bytes == lay.o_cnt + B * 4
    assert lay.o_val >= lay.read_bytes and lay.out_bytes == lay.o_val + B * k * 4
    row_off = blob[:8 * (B + 1)].view(np.int64)
    assert row_off.tolist() == np.concatenate([[0], np.cumsum(sizes)]).tolist()
    assert np.array_equal(blob[lay.o_q:lay.o_emb].view(np.float32).reshape(B, d), q.numpy())
    rows = blob[lay.o_emb:].view(np.float32).reshape(total, d)
    tiles = blob[lay.o_tiles:lay.o_tiles + 32 * lay.n_tiles].view(TILE_DTYPE)
    if cluster:
        check_cluster_table(blob[lay.o_ctas:lay.o_ctas + 32 * lay.n_ctas].view(CTA_DTYPE), sizes, row_off,
                            [dev + lay.o_emb + int(row_off[b_]) * d * 4 if sizes[b_] else 0 for b_ in range(B)], cs)
    seen = np.zeros(total, dtype=bool)
    prev = (-1, -1)
    for t in tiles:
        doc, n, r0 = int(t["doc"]), int(t["rows"]), int(t["sims_off"])
        assert 1 <= n <= lay.tile_rows and int(t["doc_rows"]) == sizes[doc] and int(t["reserved"]) == 0
        assert row_off[doc] <= r0 and r0 + n <= row_off[doc + 1]
        assert int(t["src"]) == dev + lay.o_emb + r0 * d * 4            # the tile's rows sit at sims_off in the packed matrix
        assert (doc, r0) > prev                                          # tiles of a document contiguous, ordered by sims_off
        prev = (doc, r0)
        assert not seen[r0:r0 + n].any()
        seen[r0:r0 + n] = True
    assert seen.all()
    # the packed rows ARE the documents: the oracle's scores and hits from the blob equal those from the inputs
    packed_docs = [torch.from_numpy(rows[row_off[b]:row_off[b + 1]].copy()) for b in range(B)]
    for a, b_ in zip(packed_docs, emb):
        assert torch.equal(a, b_)
    ref = R.score(emb, q)
    got = R.score(packed_docs, torch.from_numpy(blob[lay.o_q:lay.o_emb].view(np.float32).reshape(B, d).copy()))
    for x, y in zip(ref, got):
        assert torch.equal(x, y)
        assert np.array_equal(R.topk_lowest_index(x, k), R.topk_lowest_index(y, k))


def test_small_batch_argument_errors():
    lib = _lib.lib
    lay = _lib.SmallLayoutStruct()
    rows = (ctypes.c_int64 * 2)(3, -1)
    assert lib.rdv_small_batch_layout(rows, 2, 384, 5, ctypes.addressof(lay)) == _lib.E_LIMIT
    assert b"document 1" in lib.rdv_last_error()
    rows = (ctypes.c_int64 * 2)(3, 4)
    assert lib.rdv_small_batch_layout(rows, 2, 386, 5, ctypes.addressof(lay)) == _lib.E_INVALID
    assert lib.rdv_small_batch_layout(rows, 2, 384, 0, ctypes.addressof(lay)) == _lib.E_LIMIT
    assert lib.rdv_small_batch_layout(rows, 2, 384, 5, None) == _lib.E_INVALID
    assert lib.rdv_small_batch_layout(rows, 2, 384, 5, ctypes.addressof(lay)) == _lib.OK
    # buffers smaller than the layout: RDV_SMALL_GROW before anything is touched (no CUDA call on this path)
    q = np.zeros((2, 384), dtype=np.float32)
    docs = (ctypes.c_void_p * 2)(q.ctypes.data, q.ctypes.data)
    rc = lib.rdv_retrieve_small_f32(docs, rows, 2, 384, 5, q.ctypes.data, None, None, 0, None, 0, None, 0,
                                    ctypes.addressof(lay), None)
    assert rc == _lib.SMALL_GROW and lay.in_bytes > 0 and lay.out_bytes > lay.read_bytes > 0
    assert lib.rdv_small_batch_pack(docs, rows, 2, 384, q.ctypes.data, ctypes.addressof(lay), None, None) == _lib.E_INVALID


def test_retriever_small_path_host_logic_with_emulated_device(monkeypatch):
    """The Python half of Retriever._retrieve_host_small without a GPU: buffers are host tensors, the device round trip of
    rdv_retrieve_small_f32 is emulated (real layout + pack, a memmove for each copy, the oracle in place of the kernel),
    so the pointer bookkeeping, the RDV_SMALL_GROW loop, the result offsets and the list building are exercised on CPU."""
    from rag_docvqa_b200.retriever import Retriever
    import _emulated_device
    grown = _emulated_device.install(monkeypatch, small_start=True)
    base = {"compute_stats": False, "compute_stats_examples": False, "n_stats_examples": 0, "device": "cuda:0"}
    for sizes in ([30], [45, 0, 3], [60, 0, 3, 150, 1, 90, 31], [0, 0]):
        docs = len(sizes)
        for k, s, reorder in ((5, 0, False), (40, 0, False), (5, 3, True), (1, 0, False)):
            emb, q = synth.make_embeddings(sizes, 384, 311 + docs, dup_frac=0.1)
            words, boxes, labels = synth.make_words(sizes, 318 + docs, min_words=3, max_words=20, empty_chunk_every=13)
            lists = (words, boxes, labels, synth.make_images(sizes, 30, width=212, height=275, ragged_sizes=True),
                     synth.make_page_indices(sizes, 30))
            if sizes[0]:
                wide = torch.randn(sizes[0], 768, generator=torch.Generator().manual_seed(1))
                wide[:, ::2] = emb[0]
                emb[0] = wide[:, ::2]                                                    # non-contiguous rows
            retr = Retriever({**base, "chunk_num": k, "include_surroundings": s, "reorder_chunks": reorder})
            for _ in range(2):                                                           # the second call reuses the buffers
                out = retr.retrieve(emb, q, *lists)
            ref_sims = R.score([e.contiguous() for e in emb], q)
            hits = [R.topk_lowest_index(x, k) for x in out[8]]
            ref = R.gather_hits(hits, *lists, include_surroundings=s, reorder_chunks=reorder)
            for i in (0, 1, 2, 3, 4, 5, 7):
                assert out[i] == ref[i], (sizes, k, i)
            for b in range(docs):
                assert torch.equal(out[8][b], ref_sims[b]), (sizes, b)
    assert any(g != (0, 0, 0) for g in grown)                                            # RDV_SMALL_GROW was taken
