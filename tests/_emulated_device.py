"""TEST INFRASTRUCTURE: lets the Python half of the one-call host path (Retriever._retrieve_host_small ->
rdv_retrieve_small_f32) run on a machine without a GPU.  The REAL host code runs (layout, pack, pointer bookkeeping, the
RDV_SMALL_GROW loop, result offsets, list building); the two copies are memmoves and the kernel is replaced by the oracle.
Never used by the product; never used by `-m gpu` tests."""
import contextlib
import ctypes

import numpy as np
import torch

from oracle import ref_restated as R
from rag_docvqa_b200 import _lib
from rag_docvqa_b200.functional import TILE_DTYPE


def _view(ptr, n, dt):
    return np.frombuffer((ctypes.c_char * (n * np.dtype(dt).itemsize)).from_address(ptr), dtype=dt)


class EmulatedLib:
    """librdv with rdv_retrieve_small_f32 emulated on the host."""

    def __init__(self):
        self.real = _lib.lib

    def __getattr__(self, name):
        return getattr(self.real, name)

    def rdv_retrieve_small_f32(self, h_docs, rows, B, d, k, h_q, h_blob, d_blob, blob_bytes, d_out, d_out_bytes, h_out,
                               h_out_bytes, lay_p, stream):
        real = self.real
        lay = _lib.SmallLayoutStruct.from_address(lay_p)
        rc = real.rdv_small_batch_layout(rows, B, d, k, lay_p)
        if rc:
            return rc
        if lay.in_bytes > blob_bytes or lay.out_bytes > d_out_bytes or lay.read_bytes > h_out_bytes:
            return _lib.SMALL_GROW
        rc = real.rdv_small_batch_pack(h_docs, rows, B, d, h_q, lay_p, h_blob, d_blob)
        if rc:
            return rc
        ctypes.memmove(d_blob, h_blob, lay.in_bytes)                                 # "upload"
        row = _view(d_blob, B + 1, np.int64)
        q = _view(d_blob + lay.o_q, B * d, np.float32).reshape(B, d)
        tiles = _view(d_blob + lay.o_tiles, lay.n_tiles, TILE_DTYPE)
        sims = _view(d_out, max(int(row[-1]), 1), np.float32)
        idx = _view(d_out + lay.o_idx, B * k, np.int32).reshape(B, k)
        cnt = _view(d_out + lay.o_cnt, B, np.int32)
        for b in range(B):
            n = int(row[b + 1] - row[b])
            if n:
                first = [t for t in tiles if int(t["doc"]) == b][0]                  # rows of a document are contiguous
                e = _view(int(first["src"]), n * d, np.float32).reshape(n, d)
                sims[row[b]:row[b + 1]] = R.score([torch.from_numpy(e.copy())], torch.from_numpy(q[b:b + 1].copy()))[0].numpy()
            hits = R.topk_lowest_index(sims[row[b]:row[b + 1]], k)
            idx[b] = -1
            idx[b, :len(hits)] = hits
            cnt[b] = len(hits)
        ctypes.memmove(h_out, d_out, lay.read_bytes)                                 # "read-back"
        return 0


def install(monkeypatch, small_start: bool = False):
    """Patches torch.cuda stream / device handles, the Retriever's buffers (host memory) and the library (EmulatedLib).
    Returns a list that receives the buffer sizes every time the buffers are (re)allocated."""
    from rag_docvqa_b200 import functional as F
    from rag_docvqa_b200.retriever import Retriever

    class Stream:
        cuda_stream = 0

        def synchronize(self):
            pass

    monkeypatch.setattr(torch.cuda, "current_stream", lambda dev=None: Stream())
    monkeypatch.setattr(torch.cuda, "device", lambda dev: contextlib.nullcontext())
    grown = []

    def host_buffers(self, dev, n_in, n_dev, n_host):
        grown.append((n_in, n_dev, n_host))
        floor = 1 << 12 if small_start else 1 << 16                                  # small start: the grow loop runs
        c_in, c_dev, c_host = (max(2 * n, floor) for n in (n_in, n_dev, n_host))
        bufs = (torch.empty(c_in, dtype=torch.uint8), torch.empty(c_in, dtype=torch.uint8),
                torch.empty(c_dev, dtype=torch.uint8), torch.empty(c_host, dtype=torch.uint8))
        bufs = bufs + tuple(t.data_ptr() for t in bufs) + (bufs[3].numpy(),)
        self._small_bufs[dev.index] = bufs
        return bufs

    monkeypatch.setattr(Retriever, "_small_buffers", host_buffers)
    monkeypatch.setattr(F, "_lib_fn", EmulatedLib())
    return grown
