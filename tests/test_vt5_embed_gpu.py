"""GPU parity of the generator-input embeddings (csrc/vt5_embed.cu) against the reference's frozen outputs and the oracle.

Tolerance, stated: the reference computes LayerNorm + Linear in fp32 torch (error ~1e-6 relative to exact); the kernel reads
tables accumulated in fp64 and rounded once.  Both are compared with the float64 evaluation of the same formula: the kernel
must be within EMB_ATOL + EMB_RTOL * |x| of it, and within twice that of the reference's own fp32 result."""
import os

import numpy as np
import pytest
import torch

from oracle import ref_restated as R

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
EMB_RTOL, EMB_ATOL = 1e-5, 1e-5


def _weights(D, n_pos, V, n_labels, seed):
    g = torch.Generator().manual_seed(seed)
    return {
        "x_emb": torch.randn(n_pos, D, generator=g) + 0.2, "y_emb": torch.randn(n_pos, D, generator=g) * 0.7,
        "ln_weight": 1.0 + 0.2 * torch.randn(D, generator=g), "ln_bias": 0.1 * torch.randn(D, generator=g),
        "lin_weight": torch.randn(D, D, generator=g) / D ** 0.5, "lin_bias": 0.1 * torch.randn(D, generator=g),
        "shared": torch.randn(V, D, generator=g), "layout": torch.randn(n_labels, D, generator=g),
    }


def _oracle(w, eps, bbox, ids=None, labels=None, scale=1.0, dtype=torch.float64):
    c = {k: v.to(dtype) for k, v in w.items()}
    sp = R.spatial_embeddings(bbox, c["x_emb"], c["y_emb"], c["ln_weight"], c["ln_bias"], eps, c["lin_weight"], c["lin_bias"])
    if ids is None:
        return sp
    return R.vt5_input_embeds(ids, bbox, c["shared"], sp, labels, c["layout"], scale)


def _modules(w, eps, scale=1.0, layout=True):
    from rag_docvqa_b200.vt5_embed import SpatialEmbeddings, VT5InputEmbeddings
    sp = SpatialEmbeddings(w["x_emb"], w["y_emb"], w["ln_weight"], w["ln_bias"], eps, w["lin_weight"], w["lin_bias"], device=DEV)
    return sp, VT5InputEmbeddings(sp, w["shared"], w["layout"] if layout else None, scale)


def test_golden_spatial_module_and_prepare_inputs(golden_dir):
    """The frozen outputs of the reference's SpatialEmbeddings module and of VT5.prepare_inputs_for_vqa's embedding sum."""
    from rag_docvqa_b200.vt5_embed import SpatialEmbeddings, VT5InputEmbeddings
    z = np.load(os.path.join(golden_dir, "vt5_embed.npz"))
    sd = {"x_position_embeddings.weight": z["x_emb"], "y_position_embeddings.weight": z["y_emb"], "LayerNorm.weight": z["ln_weight"],
          "LayerNorm.bias": z["ln_bias"], "spatial_emb_matcher.layers.0.weight": z["lin_weight"],
          "spatial_emb_matcher.layers.0.bias": z["lin_bias"]}
    sp = SpatialEmbeddings.from_module({k: torch.from_numpy(v) for k, v in sd.items()}, device=DEV)
    got = sp(torch.from_numpy(z["bbox"]).to(DEV))
    np.testing.assert_allclose(got.cpu().numpy(), z["spatial"], rtol=2 * EMB_RTOL, atol=2 * EMB_ATOL)
    emb = VT5InputEmbeddings(sp, torch.from_numpy(z["shared"]), torch.from_numpy(z["layout"]), float(z["layout_scale"]))
    for name, labelled in (("plain", False), ("layout", True)):
        ids, boxes = torch.from_numpy(z[name + "_ids"]).to(DEV), torch.from_numpy(z[name + "_boxes"]).to(DEV)
        labels = torch.from_numpy(z[name + "_labels"]).to(DEV) if labelled else None
        out = emb(ids, boxes, labels)
        np.testing.assert_allclose(out.cpu().numpy(), z[name + "_embeds"], rtol=2 * EMB_RTOL, atol=2 * EMB_ATOL)
    emb.check()


@pytest.mark.parametrize("D,B,L", [(768, 5, 131), (512, 3, 64), (1024, 2, 77), (64, 7, 9), (100, 2, 33), (4, 1, 5), (768, 1, 1)])
def test_shapes_against_the_float64_oracle(D, B, L):
    w = _weights(D, 1024, 500, 9, seed=D + L)
    eps = 1e-12
    g = torch.Generator().manual_seed(B * 31 + L)
    bbox = torch.randint(0, 1001, (B, L, 4), generator=g)
    bbox[0, : L // 2] = torch.tensor([0, 0, 1000, 1000])             # a prompt: one box repeated
    bbox[-1, L // 2:] = 0                                            # padding
    rep = torch.rand(B, L, generator=g) < 0.4                        # tokens of one word share the box
    for b in range(B):
        for t in range(1, L):
            if rep[b, t]:
                bbox[b, t] = bbox[b, t - 1]
    ids = torch.randint(0, 500, (B, L), generator=g)
    labels = torch.randint(0, 9, (B, L), generator=g)
    sp, emb = _modules(w, eps, scale=0.7)
    for got, ref64, ref32 in (
        (sp(bbox.to(DEV)), _oracle(w, eps, bbox), _oracle(w, eps, bbox, dtype=torch.float32)),
        (emb(ids.to(DEV), bbox.to(DEV)), _oracle(w, eps, bbox, ids), _oracle(w, eps, bbox, ids, dtype=torch.float32)),
        (emb(ids.to(DEV), bbox.to(DEV), labels.to(DEV)), _oracle(w, eps, bbox, ids, labels, 0.7),
         _oracle(w, eps, bbox, ids, labels, 0.7, dtype=torch.float32)),
    ):
        assert got.shape == ref64.shape and got.dtype == torch.float32
        np.testing.assert_allclose(got.cpu().double().numpy(), ref64.numpy(), rtol=EMB_RTOL, atol=EMB_ATOL)
        np.testing.assert_allclose(got.cpu().numpy(), ref32.numpy(), rtol=2 * EMB_RTOL, atol=2 * EMB_ATOL)
        # the kernel is at least as close to the exact result as the reference's own fp32 evaluation is, up to 4x
        err_k = (got.cpu().double() - ref64).abs().max().item()
        err_r = (ref32.double() - ref64).abs().max().item()
        assert err_k <= 4 * err_r + 1e-6, (err_k, err_r)
    emb.check()
    again = emb(ids.to(DEV), bbox.to(DEV), labels.to(DEV))
    assert torch.equal(again, emb(ids.to(DEV), bbox.to(DEV), labels.to(DEV)))     # run-to-run bit-identical


def test_trimmed_views_of_the_gather_buffers_and_empty_batches():
    """PackedInputs are (B, max_len) buffers cut to the longest row: views with a row pitch go in without a copy; views with
    unrelated pitches take dense copies; B = 0 and L = 0 launch nothing."""
    w = _weights(128, 1024, 300, 5, seed=3)
    eps = 1e-12
    sp, emb = _modules(w, eps, scale=1.3)
    g = torch.Generator().manual_seed(9)
    B, max_len, L = 4, 64, 41
    ids_buf = torch.randint(0, 300, (B, max_len), generator=g).to(DEV)
    box_buf = torch.randint(0, 1001, (B, max_len, 4), generator=g).to(DEV)
    lab_buf = torch.randint(0, 5, (B, max_len), generator=g).to(DEV)
    ids, boxes, labels = ids_buf[:, :L], box_buf[:, :L], lab_buf[:, :L]
    ref = _oracle(w, eps, boxes.cpu(), ids.cpu(), labels.cpu(), 1.3)
    np.testing.assert_allclose(emb(ids, boxes, labels).cpu().double().numpy(), ref.numpy(), rtol=EMB_RTOL, atol=EMB_ATOL)
    odd = torch.randint(0, 300, (B, max_len + 3), generator=g).to(DEV)[:, :L]            # another pitch
    ref2 = _oracle(w, eps, boxes.cpu(), odd.cpu())
    np.testing.assert_allclose(emb(odd, boxes).cpu().double().numpy(), ref2.numpy(), rtol=EMB_RTOL, atol=EMB_ATOL)
    tr = boxes.transpose(0, 1).contiguous().transpose(0, 1)                              # not row-major at all
    np.testing.assert_allclose(sp(tr).cpu().double().numpy(), _oracle(w, eps, boxes.cpu()).numpy(), rtol=EMB_RTOL, atol=EMB_ATOL)
    assert emb(ids[:0], boxes[:0]).shape == (0, L, 128)
    assert emb(ids[:, :0], boxes[:, :0]).shape == (B, 0, 128)
    emb.check()


def test_out_of_range_indices_are_flagged_not_read():
    """torch.nn.Embedding raises on an index outside the table; the kernel clamps it (no out-of-bounds read) and check()
    raises afterwards, naming the table."""
    w = _weights(64, 256, 50, 3, seed=4)
    sp, emb = _modules(w, 1e-12)
    ok_ids = torch.zeros((1, 8), dtype=torch.int64, device=DEV)
    ok_box = torch.zeros((1, 8, 4), dtype=torch.int64, device=DEV)
    ok_lab = torch.zeros((1, 8), dtype=torch.int64, device=DEV)
    emb(ok_ids, ok_box, ok_lab)
    emb.check()
    for what, ids, box, lab in (("box coordinate", ok_ids, ok_box.clone().index_fill_(1, torch.tensor([3], device=DEV), 256), ok_lab),
                                ("box coordinate", ok_ids, ok_box.clone().index_fill_(1, torch.tensor([7], device=DEV), -1), ok_lab),
                                ("token id", ok_ids.clone().fill_(50), ok_box, ok_lab),
                                ("layout label", ok_ids, ok_box, ok_lab.clone().fill_(-2))):
        out = emb(ids, box, lab)
        assert torch.isfinite(out).all()
        with pytest.raises(IndexError, match=what):
            emb.check()
    emb.check()                                                                         # the flag was cleared
    with pytest.raises(ValueError):
        emb(ok_ids, ok_box.to(torch.int32))
    with pytest.raises(ValueError):
        _modules(w, 1e-12, layout=False)[1](ok_ids, ok_box, ok_lab)
    with pytest.raises(RuntimeError):
        emb(ok_ids.cpu(), ok_box)


def test_after_the_gather_kernel():
    """ids / boxes / labels straight from retrieve_packed (views of the gather's (B, max_len) buffers): the embeddings equal
    the oracle's on the same tensors (that the tensors are the reference's: tests/test_retriever_gpu.py)."""
    import zlib
    from rag_docvqa_b200 import synth
    from rag_docvqa_b200.docstore import DocStore
    from rag_docvqa_b200.retriever import Retriever
    batch = synth.make_text_batch("C2", with_lists=True, docs=6, seed=21, dup_frac=0.0)
    words, boxes, labels = batch["words_text_chunks"], batch["words_box_chunks"], batch["layout_labels_chunks"]
    table = synth.make_tokens_for_words(words, seed=9)
    store = DocStore.from_lists(words, boxes, labels, batch["page_indices"], lambda w: table.get(w, [2]), torch.device(DEV),
                                images=batch["images"])
    prompts = [[5 + (zlib.crc32(t.encode()) % 1000) for t in ("question: what is item %d about ?  context: " % b).split()]
               for b in range(len(words))]
    retr = Retriever({"chunk_num": 5, "include_surroundings": 0, "compute_stats": False, "compute_stats_examples": False,
                      "n_stats_examples": 0})
    packed, _ = retr.retrieve_packed([e.to(DEV) for e in batch["text_embeddings"]], batch["question_embeddings"].to(DEV),
                                     store, prompts, max_source_length=512, with_layout_labels=True)
    V = int(packed.input_ids.max().item()) + 1
    n_labels = int(packed.layout_labels.max().item()) + 1
    w = _weights(96, 1024, V, n_labels, seed=5)
    _, emb = _modules(w, 1e-12, scale=0.5)
    out = emb(packed.input_ids, packed.boxes, packed.layout_labels)
    emb.check()
    assert out.shape == (len(words), packed.longest, 96)
    ref = _oracle(w, 1e-12, packed.boxes.cpu(), packed.input_ids.cpu(), packed.layout_labels.cpu(), 0.5)
    np.testing.assert_allclose(out.cpu().double().numpy(), ref.numpy(), rtol=EMB_RTOL, atol=EMB_ATOL)
    # prepare_inputs: the reference's return values (src/VT5.py:205-207), the text part written with the row pitch of the
    # concatenated buffer -- bit-identical to the contiguous call, the visual tokens behind it, the masks concatenated
    g = torch.Generator().manual_seed(6)
    vis = torch.randn(len(words), 7, 96, generator=g).to(DEV)
    vmask = torch.ones(len(words), 7, dtype=torch.int64, device=DEV)
    vmask[0, 5:] = 0
    embeds, mask = emb.prepare_inputs(packed, vis, vmask)
    emb.check()
    assert embeds.shape == (len(words), packed.longest + 7, 96)
    assert torch.equal(embeds[:, :packed.longest], out) and torch.equal(embeds[:, packed.longest:], vis)
    assert torch.equal(mask, torch.cat([packed.attention_mask, vmask], dim=1))
    plain, mask0 = emb.prepare_inputs(packed)
    assert torch.equal(plain, out) and torch.equal(mask0, packed.attention_mask)
