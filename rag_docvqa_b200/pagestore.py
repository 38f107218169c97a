"""Page images on the device + the gather of the retrieved patches into the generator's visual input.

Host side: the documents' PIL pages (src/RAGVT5.py:208-224 `images`) are converted ONCE per batch of documents
into tightly packed RGB uint8 rows in one device buffer (identical page objects are stored once).  Device side:
rdv_visual_pack (csrc/visual_pack.cu) turns the gather kernel's crop rectangles into the S x S image the reference
builds with page.crop -> concatenate_patches(mode="grid") -> feature-extractor resize (src/_modules.py:2102-2121,
src/utils.py:180-231, src/_modules.py:133), bit-exact against Pillow, plus the normalised fp32 pixel_values.
"""
from __future__ import annotations

import ctypes
import math
from typing import NamedTuple, Optional, Sequence

import numpy as np
import torch

from . import _lib
from .functional import _stream_ptr

BILINEAR, BICUBIC = 2, 3        # PIL.Image.Resampling values


class VisualInputs(NamedTuple):
    image_u8: torch.Tensor        # (B, S, S, 3) uint8: resize(concatenate_patches(crops, "grid"))
    pixel_values: Optional[torch.Tensor]   # (B, 3, S, S) fp32 = (u8 / 255 - mean) / std
    status: torch.Tensor          # (B,) int32: 0 ok, 1 capacity, 2 degenerate canvas (the reference raises there)


class PageStore:
    """RGB pages of a batch of documents on the device (see `rdv_pagestore` in include/rdv.h)."""

    def __init__(self, images: Sequence[Sequence["PIL.Image.Image"]], device):
        self.B = len(images)
        self.device = device
        doc_page_off = np.zeros(self.B + 1, dtype=np.int32)
        np.cumsum([len(p) for p in images], out=doc_page_off[1:])
        P = int(doc_page_off[-1])
        wh = np.zeros((P, 2), dtype=np.int32)
        page_off = np.zeros(P, dtype=np.int64)
        unique, chunks, total = {}, [], 0
        i = 0
        for pages in images:
            for im in pages:
                key = id(im)
                if key not in unique:
                    arr = np.asarray(im.convert("RGB") if im.mode != "RGB" else im, dtype=np.uint8)
                    arr = np.ascontiguousarray(arr)
                    unique[key] = (total, arr.shape[1], arr.shape[0])
                    chunks.append(arr.reshape(-1))
                    total += arr.size
                    total = (total + 15) // 16 * 16
                off, w, h = unique[key]
                page_off[i] = off
                wh[i] = (w, h)
                i += 1
        blob = torch.empty(max(total, 16), dtype=torch.uint8, pin_memory=True)
        raw = blob.numpy()
        pos = 0
        for arr in chunks:
            raw[pos:pos + arr.size] = arr
            pos = (pos + arr.size + 15) // 16 * 16
        self.pixels = blob.to(device, non_blocking=True)
        meta = np.concatenate([doc_page_off.view(np.uint8), np.zeros((-doc_page_off.nbytes) % 16, np.uint8),
                               wh.reshape(-1).view(np.uint8), np.zeros((-wh.nbytes) % 16, np.uint8), page_off.view(np.uint8)])
        self.meta = torch.from_numpy(meta).pin_memory().to(device, non_blocking=True)
        o_wh = doc_page_off.nbytes + (-doc_page_off.nbytes) % 16
        o_off = o_wh + wh.nbytes + (-wh.nbytes) % 16
        self.struct = _lib.PageStoreStruct()
        self.struct.B = self.B
        base = self.meta.data_ptr()
        self.struct.doc_page_off = base
        self.struct.page_wh = base + o_wh
        self.struct.page_off = base + o_off
        self.struct.pixels = self.pixels.data_ptr()
        self.doc_page_off_host = doc_page_off
        self.max_w = int(wh[:, 0].max()) if P else 1
        self.max_h = int(wh[:, 1].max()) if P else 1
        self.n_pages, self.n_unique, self.bytes = P, len(unique), total

    @classmethod
    def from_images(cls, images, device) -> "PageStore":
        return cls(images, device)

    def prepare_pack(self, hit_page: torch.Tensor, hit_rect: torch.Tensor, hit_cnt: torch.Tensor, out_size: int = 224,
                     resample: int = BICUBIC, mean=(0.5, 0.5, 0.5), std=(0.5, 0.5, 0.5),
                     with_pixel_values: bool = True) -> "VisualPlan":
        """Allocates workspaces / outputs and fills the argument block of rdv_visual_pack (no launch)."""
        dev = self.device
        B, k = hit_page.shape
        if B != self.B:
            raise ValueError("visual pack: %d documents of hits, store has %d" % (B, self.B))
        support = 1.0 if resample == BILINEAR else 2.0
        cap_h = 2 * math.ceil(support * max(self.max_w / out_size, 1.0)) + 1
        rows_cap = max(k * self.max_h, 5)
        cap_v = 2 * math.ceil(support * max(rows_cap / out_size, 1.0)) + 1
        t = dict(hit_page=hit_page, hit_rect=hit_rect, hit_cnt=hit_cnt,
                 layout=torch.empty((B, 8 + 4 * k), dtype=torch.int32, device=dev),
                 coeff_h=torch.empty((B, out_size, cap_h + 2), dtype=torch.int32, device=dev),
                 coeff_v=torch.empty((B, out_size, cap_v + 2), dtype=torch.int32, device=dev),
                 temp=torch.empty((B, rows_cap, out_size, 3), dtype=torch.uint8, device=dev),
                 out_u8=torch.empty((B, out_size, out_size, 3), dtype=torch.uint8, device=dev),
                 out_px=torch.empty((B, 3, out_size, out_size), dtype=torch.float32, device=dev) if with_pixel_values else None,
                 status=torch.empty((B,), dtype=torch.int32, device=dev))
        a = _lib.VisualArgsStruct()
        a.hit_page, a.hit_rect, a.hit_cnt = hit_page.data_ptr(), hit_rect.data_ptr(), hit_cnt.data_ptr()
        a.k, a.out_size, a.filter = k, out_size, int(resample)
        a.ksize_cap_h, a.ksize_cap_v, a.rows_cap = cap_h, cap_v, rows_cap
        a.max_page_w = self.max_w
        for c in range(3):
            a.mean[c], a.std[c] = float(mean[c]), float(std[c])
        a.layout, a.coeff_h, a.coeff_v = t["layout"].data_ptr(), t["coeff_h"].data_ptr(), t["coeff_v"].data_ptr()
        a.temp, a.out_u8, a.status = t["temp"].data_ptr(), t["out_u8"].data_ptr(), t["status"].data_ptr()
        a.out_px = t["out_px"].data_ptr() if t["out_px"] is not None else None
        return VisualPlan(self, a, t)

    def pack(self, hit_page, hit_rect, hit_cnt, **options) -> VisualInputs:
        plan = self.prepare_pack(hit_page, hit_rect, hit_cnt, **options)
        plan.launch()
        return plan.result()


class VisualPlan:
    """A filled argument block of rdv_visual_pack: launch() enqueues the three kernels (no sync)."""

    def __init__(self, store: PageStore, args, tensors: dict):
        self.store, self.args, self.t = store, args, tensors
        self._ps_ref, self._args_ref = ctypes.byref(store.struct), ctypes.byref(args)

    def launch(self, stream: Optional[int] = None) -> None:
        rc = _lib.lib.rdv_visual_pack(self._ps_ref, self._args_ref,
                                      _stream_ptr(self.store.device) if stream is None else stream)
        if rc:
            _lib.check(rc)

    def result(self) -> VisualInputs:
        return VisualInputs(self.t["out_u8"], self.t["out_px"], self.t["status"])


# ---- Pix2Struct flattened patches (SURVEY.md section 8 row a12, Pix2Struct half) -------------------------------------
P2S_IMG_DTYPE = np.dtype([("doc", "<i4"), ("page", "<i4"), ("x0", "<i4"), ("y0", "<i4"), ("x1", "<i4"), ("y1", "<i4"),
                          ("rows", "<i4"), ("cols", "<i4"), ("kept", "<i4"), ("out_start", "<i4"), ("row_offset", "<i4"),
                          ("reserved", "<i4"), ("temp_off", "<i8")])      # rdv_p2s_img, 56 bytes


class Pix2StructInputs(NamedTuple):
    flattened_patches: torch.Tensor     # (B, max_total_patches, 2 + patch*patch*3) fp32
    attention_mask: torch.Tensor        # (B, max_total_patches) fp32


def plan_pix2struct(crops: Sequence[Sequence[Sequence[int]]], doc_page_off: np.ndarray, max_total_patches: int, patch: int):
    """The per-image plan of extract_multi_image_flattened_patches (src/custom_pix2struct_processor.py:97-132, :52-57):
    patch budget per image, patch grid from the budget (Python float arithmetic, as the reference), output offsets,
    row-id offsets.  crops[b] = [(page_in_doc, x0, y0, x1, y1), ...]."""
    recs, doc_total, temp_off = [], [], 0
    for b, doc in enumerate(crops):
        if len(doc) == 0:
            raise ValueError("No images provided.")                                  # :109
        per = max_total_patches // len(doc)                                           # :110
        out_start, row_offset = 0, 0
        for (p, x0, y0, x1, y1) in doc:
            w, h = int(x1) - int(x0), int(y1) - int(y0)
            if w <= 0 or h <= 0:
                raise ValueError("document %d: empty crop %r" % (b, (x0, y0, x1, y1)))
            scale = math.sqrt(per * (patch / h) * (patch / w))                        # :52
            rows = max(min(math.floor(scale * h / patch), per), 1)                    # :53
            cols = max(min(math.floor(scale * w / patch), per), 1)                    # :54
            kept = min(rows * cols, per)
            recs.append((b, int(doc_page_off[b]) + int(p), int(x0), int(y0), int(x1), int(y1), rows, cols, kept, out_start,
                         row_offset, 0, temp_off))
            out_start += kept
            row_offset += rows                                                        # int(row_ids.max()) (:95)
            temp_off += h * cols * patch * 3
        doc_total.append(out_start)
    return np.array(recs, dtype=P2S_IMG_DTYPE), np.asarray(doc_total, dtype=np.int32), temp_off


def _pack_pix2struct(self, crops, max_total_patches: int = 2048, patch: int = 16, normalize: bool = True) -> Pix2StructInputs:
    """Crops of the store's pages -> (flattened_patches, attention_mask) on the device (rdv_pix2struct_patches).
    crops[b] = [(page_in_doc, x0, y0, x1, y1), ...] in the order the reference would pass the images."""
    dev = self.device
    if len(crops) != self.B:
        raise ValueError("pix2struct pack: %d documents of crops, store has %d" % (len(crops), self.B))
    images, doc_total, temp_floats = plan_pix2struct(crops, self.doc_page_off_host, max_total_patches, patch)
    depth = 2 + patch * patch * 3
    blob = torch.empty(images.nbytes + doc_total.nbytes + 16, dtype=torch.uint8, pin_memory=True)
    raw = blob.numpy()
    raw[:images.nbytes] = images.view(np.uint8).reshape(-1)
    o_tot = (images.nbytes + 15) // 16 * 16
    raw[o_tot:o_tot + doc_total.nbytes] = doc_total.view(np.uint8)
    small = blob.to(dev, non_blocking=True)
    t = dict(small=small, stats=torch.empty((max(len(images), 1), 2), dtype=torch.int64, device=dev),   # byte sums
             temp=torch.empty((max(temp_floats, 1),), dtype=torch.float32, device=dev),
             out=torch.empty((self.B, max_total_patches, depth), dtype=torch.float32, device=dev),
             mask=torch.empty((self.B, max_total_patches), dtype=torch.float32, device=dev))
    a = _lib.P2SArgsStruct()
    a.images, a.n_images, a.n_docs = small.data_ptr(), len(images), self.B
    a.max_total, a.patch, a.do_normalize = max_total_patches, patch, 1 if normalize else 0
    a.max_rw = int((images["cols"] * patch).max()) if len(images) else 0
    a.max_rwh = int((images["cols"].astype(np.int64) * images["rows"] * patch * patch).max()) if len(images) else 0
    a.max_h = int((images["y1"] - images["y0"]).max()) if len(images) else 0
    a.stats, a.temp, a.doc_total = t["stats"].data_ptr(), t["temp"].data_ptr(), small.data_ptr() + o_tot
    a.out, a.mask = t["out"].data_ptr(), t["mask"].data_ptr()
    with torch.cuda.device(dev):
        _lib.check(_lib.lib.rdv_pix2struct_patches(ctypes.byref(self.struct), ctypes.byref(a), _stream_ptr(dev)))
    self._p2s_keepalive = t
    return Pix2StructInputs(t["out"], t["mask"])


PageStore.pack_pix2struct = _pack_pix2struct
