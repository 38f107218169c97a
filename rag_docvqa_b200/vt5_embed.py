"""The generator's input embeddings on the device: what VT5.prepare_inputs_for_vqa does after it has built
input_ids / boxes / layout labels (reference src/VT5.py:194-204), fed straight from the gather kernel's tensors.

    spatial = SpatialEmbeddings.from_module(vt5.spatial_embedding)            # tables built once per model
    embed = VT5InputEmbeddings.from_model(vt5)                                 # + shared / layout tables
    packed, res = retriever.retrieve_packed(...)                               # ids / boxes / mask on the device
    input_embeds = embed(packed.input_ids, packed.boxes, packed.layout_labels) # (B, longest, D), one launch

`SpatialEmbeddings` mirrors the reference module of the same name (src/_modules.py:48-86; inference only: dropout is the
identity, nothing is trainable here) -- `spatial(bbox)` returns what the module's forward returns.  The LayerNorm and the
Linear of the reference are folded into per-coordinate tables (csrc/vt5_embed.cu explains the algebra); results agree with
the reference's fp32 modules to fp32 rounding (tests/test_vt5_embed_gpu.py states the tolerance).  CUDA only.
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib
from .functional import _f32_contig_aligned, _require_cuda, _stream_ptr

_lib_fn = _lib.lib


class SpatialEmbeddings:
    """Drop-in for the inference forward of src/_modules.py:48-86."""

    def __init__(self, x_position_embeddings: torch.Tensor, y_position_embeddings: torch.Tensor, ln_weight: torch.Tensor,
                 ln_bias: torch.Tensor, eps: float, lin_weight: torch.Tensor, lin_bias: torch.Tensor = None, device=None):
        device = torch.device(device if device is not None else x_position_embeddings.device)
        if device.type != "cuda":
            raise RuntimeError("SpatialEmbeddings: a CUDA device is required: rag_docvqa_b200 has no CPU fallback")
        x = _f32_contig_aligned(x_position_embeddings.detach().to(device))
        y = _f32_contig_aligned(y_position_embeddings.detach().to(device))
        if x.dim() != 2 or x.shape != y.shape:
            raise ValueError("SpatialEmbeddings: x / y position tables must both be (max_2d_position_embeddings, hidden)")
        n_pos, D = int(x.shape[0]), int(x.shape[1])
        w = _f32_contig_aligned(lin_weight.detach().to(device))
        if tuple(w.shape) != (D, D):
            raise ValueError("SpatialEmbeddings: spatial_emb_matcher must be one Linear(hidden, hidden) (src/_modules.py:66)")
        g = _f32_contig_aligned(ln_weight.detach().to(device))
        b = _f32_contig_aligned(ln_bias.detach().to(device))
        lb = None if lin_bias is None else _f32_contig_aligned(lin_bias.detach().to(device))
        self.device, self.D, self.n_pos, self.eps = device, D, n_pos, float(eps)
        self.xw = torch.empty((n_pos, D), dtype=torch.float32, device=device)
        self.yw = torch.empty((n_pos, D), dtype=torch.float32, device=device)
        self.gxx = torch.empty((n_pos, n_pos), dtype=torch.float32, device=device)
        self.gxy = torch.empty((n_pos, n_pos), dtype=torch.float32, device=device)
        self.gyy = torch.empty((n_pos, n_pos), dtype=torch.float32, device=device)
        self.c = torch.empty((D,), dtype=torch.float32, device=device)
        self.bad = torch.zeros((1,), dtype=torch.int32, device=device)
        self._work = {}                                                 # stream -> the kernel's chunk counter (kept zero)
        with torch.cuda.device(device):
            means = torch.empty((2 * n_pos,), dtype=torch.float64, device=device)
            _lib.check(_lib_fn.rdv_vt5_embed_tables_build(
                x.data_ptr(), y.data_ptr(), n_pos, D, g.data_ptr(), b.data_ptr(), w.data_ptr(),
                0 if lb is None else lb.data_ptr(), means.data_ptr(), self.xw.data_ptr(), self.yw.data_ptr(),
                self.gxx.data_ptr(), self.gxy.data_ptr(), self.gyy.data_ptr(), self.c.data_ptr(), _stream_ptr(device)))
            torch.cuda.current_stream(device).synchronize()           # x / y / w / means may be freed after this returns
        self._struct = self._make_struct()

    @classmethod
    def from_module(cls, module, device=None) -> "SpatialEmbeddings":
        """From the reference's module (or its state dict): x_position_embeddings, y_position_embeddings, LayerNorm,
        spatial_emb_matcher.layers[0] (src/_modules.py:56-66)."""
        sd = module if isinstance(module, dict) else module.state_dict()
        eps = 1e-12 if isinstance(module, dict) else float(module.LayerNorm.eps)
        return cls(sd["x_position_embeddings.weight"], sd["y_position_embeddings.weight"], sd["LayerNorm.weight"],
                   sd["LayerNorm.bias"], eps, sd["spatial_emb_matcher.layers.0.weight"],
                   sd.get("spatial_emb_matcher.layers.0.bias"), device=device)

    def _make_struct(self, shared=None, layout=None, layout_scale: float = 1.0) -> _lib.EmbedTablesStruct:
        t = _lib.EmbedTablesStruct()
        t.D, t.n_pos, t.eps = self.D, self.n_pos, self.eps
        t.xw, t.yw, t.gxx, t.gxy, t.gyy, t.c = (x.data_ptr() for x in (self.xw, self.yw, self.gxx, self.gxy, self.gyy, self.c))
        t.V, t.shared = (0, None) if shared is None else (int(shared.shape[0]), shared.data_ptr())
        t.n_labels, t.layout = (0, None) if layout is None else (int(layout.shape[0]), layout.data_ptr())
        t.layout_scale = float(layout_scale)
        return t

    def _launch(self, struct, ids, boxes, labels, extra_rows: int = 0) -> torch.Tensor:
        _require_cuda(boxes, "boxes")
        if boxes.dim() != 3 or boxes.shape[2] != 4 or boxes.dtype != torch.int64:
            raise ValueError("boxes must be (B, L, 4) int64 (tensor_boxes of src/VT5.py:174), got %s %s" % (tuple(boxes.shape), boxes.dtype))
        B, L = int(boxes.shape[0]), int(boxes.shape[1])
        ld = L
        if B and L:
            # the gather's (B, max_len) buffers trimmed to the longest row are views with a row pitch: no copy for those
            pitched = boxes.stride(2) == 1 and boxes.stride(1) == 4 and boxes.stride(0) % 4 == 0 and boxes.stride(0) >= 4 * L
            if not pitched or boxes.data_ptr() % 16:
                boxes = boxes.contiguous()
            ld = boxes.stride(0) // 4 if B > 1 else L
        # extra_rows: room after every row's L tokens (the generator's visual tokens): the kernel writes with that row pitch
        out = torch.empty((B, L + int(extra_rows), self.D), dtype=torch.float32, device=boxes.device)

        def rows(t, what):
            if t is None:
                return None
            _require_cuda(t, what)
            if tuple(t.shape) != (B, L) or t.dtype != torch.int64:
                raise ValueError("%s must be (B, L) int64 matching boxes, got %s %s" % (what, tuple(t.shape), t.dtype))
            if B and L and (t.stride(1) != 1 or (B > 1 and t.stride(0) != ld)):
                if ld != L:                                             # one pitch for all three: fall back to dense copies
                    return "dense"
                t = t.contiguous()
            return t
        ids_r, lab_r = rows(ids, "input_ids"), rows(labels, "layout_labels")
        if isinstance(ids_r, str) or isinstance(lab_r, str):
            boxes, ld = boxes.contiguous(), L
            ids_r = None if ids is None else ids.contiguous()
            lab_r = None if labels is None else labels.contiguous()
        with torch.cuda.device(boxes.device):
            stream = _stream_ptr(boxes.device)
            work = self._work.get(stream)
            if work is None:
                work = self._work[stream] = torch.zeros((2,), dtype=torch.int32, device=self.device)
            _lib.check(_lib_fn.rdv_vt5_input_embeds_f32(
                ctypes.byref(struct), 0 if ids_r is None else ids_r.data_ptr(), boxes.data_ptr(),
                0 if lab_r is None else lab_r.data_ptr(), B, L, ld, out.data_ptr(), L + int(extra_rows), self.bad.data_ptr(),
                work.data_ptr(), stream))
        return out

    def forward(self, bbox: torch.Tensor) -> torch.Tensor:
        """(B, L, 4) int64 boxes in [0, max_2d_position_embeddings) -> (B, L, hidden), as src/_modules.py:70-86."""
        return self._launch(self._struct, None, bbox, None)

    __call__ = forward

    def check(self) -> None:
        """Raises IndexError if a launch since the last check saw an index outside its table (torch.nn.Embedding raises at
        the call; here the entry was clamped and flagged).  Synchronises."""
        bad = int(self.bad.item())
        if bad:
            self.bad.zero_()
            what = [n for bit, n in ((1, "box coordinate"), (2, "token id"), (4, "layout label")) if bad & bit]
            raise IndexError("index out of range in %s table(s)" % " / ".join(what))


class VT5InputEmbeddings:
    """semantic + spatial (+ layout * scale): the embedding sum of VT5.prepare_inputs_for_vqa (src/VT5.py:194-204), one
    launch over the gather's tensors.  The visual tokens the reference concatenates after it (:205) come from the
    generator's own ViT and are not computed here; `prepare_inputs` takes them and returns what the reference's
    prepare_inputs_for_vqa returns, writing the text part straight into the concatenated buffer."""

    def __init__(self, spatial: SpatialEmbeddings, shared_weight: torch.Tensor, layout_weight: torch.Tensor = None,
                 layout_scale: float = 1.0):
        dev = spatial.device
        self.spatial = spatial
        self.shared = _f32_contig_aligned(shared_weight.detach().to(dev))
        self.layout = None if layout_weight is None else _f32_contig_aligned(layout_weight.detach().to(dev))
        if self.shared.dim() != 2 or self.shared.shape[1] != spatial.D or (self.layout is not None and self.layout.shape[1] != spatial.D):
            raise ValueError("VT5InputEmbeddings: the token / layout tables must be (rows, %d)" % spatial.D)
        self.layout_scale = float(layout_scale)
        self._plain = spatial._make_struct(self.shared)
        self._with_layout = None if self.layout is None else spatial._make_struct(self.shared, self.layout, self.layout_scale)

    @classmethod
    def from_model(cls, model, device=None) -> "VT5InputEmbeddings":
        """From a VT5ForConditionalGeneration (src/VT5.py:17-38): language_backbone.shared, spatial_embedding and, when the
        model embeds layout labels, layout_embedding / layout_embedding_scale."""
        spatial = SpatialEmbeddings.from_module(model.spatial_embedding, device=device)
        layout = getattr(model, "layout_embedding", None)
        scale = getattr(model, "layout_embedding_scale", 1.0)
        return cls(spatial, model.language_backbone.shared.weight, None if layout is None else layout.weight, float(scale))

    def __call__(self, input_ids: torch.Tensor, boxes: torch.Tensor, layout_labels: torch.Tensor = None) -> torch.Tensor:
        if layout_labels is not None and self._with_layout is None:
            raise ValueError("layout labels given, but the model has no layout embedding (use_layout_labels != 'Embed')")
        _require_cuda(input_ids, "input_ids")
        return self.spatial._launch(self._plain if layout_labels is None else self._with_layout, input_ids, boxes, layout_labels)

    def prepare_inputs(self, packed, visual_embedding: torch.Tensor = None, visual_mask: torch.Tensor = None):
        """(input_embeds, attention_mask) of VT5.prepare_inputs_for_vqa (src/VT5.py:194-207) from the gather's PackedInputs
        (docstore.PackedInputs: input_ids / boxes / attention_mask / layout_labels on the device) and the generator's
        visual tokens (B, n_visual, hidden) + their mask (B, n_visual): the embedding kernel writes rows [0, longest) of the
        (B, longest + n_visual, hidden) result directly, so the reference's torch.cat of the text part costs nothing."""
        labels = packed.layout_labels if self._with_layout is not None else None
        n_vis = 0 if visual_embedding is None else int(visual_embedding.shape[1])
        if visual_embedding is not None:
            _require_cuda(visual_embedding, "visual_embedding")
            if visual_embedding.dim() != 3 or visual_embedding.shape[0] != packed.input_ids.shape[0] or visual_embedding.shape[2] != self.spatial.D:
                raise ValueError("visual_embedding must be (B, n_visual, %d)" % self.spatial.D)
        _require_cuda(packed.input_ids, "input_ids")
        struct = self._plain if labels is None else self._with_layout
        embeds = self.spatial._launch(struct, packed.input_ids, packed.boxes, labels, extra_rows=n_vis)
        mask = packed.attention_mask
        if n_vis:
            L = int(packed.input_ids.shape[1])
            embeds[:, L:].copy_(visual_embedding)
            if visual_mask is None:
                visual_mask = torch.ones(visual_embedding.shape[:2], dtype=mask.dtype, device=mask.device)
            mask = torch.cat([mask, visual_mask.to(mask.dtype)], dim=1)
        return embeds, mask

    def check(self) -> None:
        self.spatial.check()
