"""TEST INFRASTRUCTURE ONLY -- freezes the generator-input embeddings as the UNMODIFIED reference computes them: the
reference's SpatialEmbeddings module (src/_modules.py:48-86, eval mode) with seeded weights, and the embedding sum of
VT5.prepare_inputs_for_vqa (src/VT5.py:141-206) run as written on a stand-in `self` (no tokenizer / checkpoint exists
offline: a table tokenizer, nn.Embedding tables with seeded weights) -> tests/golden/vt5_embed.npz.

    PYTHONDONTWRITEBYTECODE=1 python -m oracle.make_golden_embed        (build container: /root/reference must exist)
"""
import importlib
import json
import os
import sys
import types
import zlib

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.ref_import import import_reference  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
D, N_POS, VOCAB, N_LABELS = 96, 1024, 1200, 12


def main():
    modules, utils, _ = import_reference()
    vt5 = importlib.import_module("src.VT5")
    torch.manual_seed(20260)
    cfg = types.SimpleNamespace(max_2d_position_embeddings=N_POS, hidden_size=D, layer_norm_eps=1e-12, hidden_dropout_prob=0.1)
    spatial = modules.SpatialEmbeddings(cfg).eval()                       # the reference module, unmodified
    with torch.no_grad():                                                 # trained-looking LayerNorm parameters
        spatial.LayerNorm.weight.copy_(1.0 + 0.2 * torch.randn(D))
        spatial.LayerNorm.bias.copy_(0.1 * torch.randn(D))
        spatial.x_position_embeddings.weight.add_(0.3)                    # rows with a non-zero mean
    shared = torch.nn.Embedding(VOCAB, D)
    layout = torch.nn.Embedding(N_LABELS, D)
    scale = 0.7

    # 1. the module alone on boxes that cover the corners of the table
    g = torch.Generator().manual_seed(5)
    bbox = torch.randint(0, 1001, (3, 37, 4), generator=g)
    bbox[0, 0] = torch.tensor([0, 0, 0, 0])
    bbox[0, 1] = torch.tensor([0, 0, 1000, 1000])
    bbox[0, 2] = torch.tensor([N_POS - 1, N_POS - 1, N_POS - 1, N_POS - 1])
    bbox[1, 5:9] = bbox[1, 4]                                             # a word of five tokens
    with torch.no_grad():
        sp = spatial(bbox)

    # 2. prepare_inputs_for_vqa as written, on words / boxes / labels
    words = [["alpha", "beta", "gamma", "delta"], ["omega"], []]
    boxes = [[[0.1, 0.2, 0.3, 0.4], [0.5, 0.5, 0.75, 0.625], [0.0, 0.0, 1.0, 1.0], [0.999, 0.001, 0.9999, 0.5]],
             [[0.25, 0.125, 0.5, 0.875]], []]
    labels = [[1, 2, 3, 11], [0], []]
    table = {"alpha": [11, 12], "beta": [13], "gamma": [14, 15, 16], "delta": [17], "omega": [18, 19]}

    class FakeTokenizer:
        eos_token_id, pad_token_id = 1, 0

        def __call__(self, text, **kw):
            if text.startswith("question: "):
                ids = [20 + (zlib.crc32(t.encode()) % 1000) for t in text.split()]
            else:
                ids = list(table.get(text, [2]))
            return types.SimpleNamespace(input_ids=ids + [self.eos_token_id])

    out = {}
    for name, use_layout in (("plain", "Default"), ("layout", "Embed")):
        fake_self = types.SimpleNamespace(
            tokenizer=FakeTokenizer(), max_source_length=24, use_layout_labels=use_layout,
            language_backbone=types.SimpleNamespace(device="cpu", shared=shared),
            spatial_embedding=spatial, layout_embedding=layout, layout_embedding_scale=torch.tensor(scale),
            visual_embedding=lambda ims: (torch.zeros(len(ims), 0, D), torch.zeros(len(ims), 0, dtype=torch.long)))
        captured = {}

        def spy(bx, _c=captured):
            _c["boxes"] = bx.clone()
            return spatial(bx)
        fake_self.spatial_embedding = spy
        questions = ["what is item %d ?" % b for b in range(len(words))]
        with torch.no_grad():
            embeds, attn, _, lab = vt5.VT5ForConditionalGeneration.prepare_inputs_for_vqa(
                fake_self, questions, words, boxes, labels, [None] * len(words), None)
            ids, _, _, _ = vt5.VT5ForConditionalGeneration.prepare_inputs_for_vqa(
                fake_self, questions, words, boxes, labels, [None] * len(words), None, return_ids=True)
        out[name + "_embeds"] = embeds.numpy()
        out[name + "_ids"] = ids.numpy()
        out[name + "_boxes"] = captured["boxes"].numpy()
        out[name + "_mask"] = attn.numpy()
        if lab is not None:
            out[name + "_labels"] = lab.numpy()
    sd = {k: v.detach().numpy() for k, v in spatial.state_dict().items()}
    np.savez_compressed(
        os.path.join(GOLDEN, "vt5_embed.npz"), bbox=bbox.numpy(), spatial=sp.numpy(), shared=shared.weight.detach().numpy(),
        layout=layout.weight.detach().numpy(), layout_scale=np.float32(scale), eps=np.float64(cfg.layer_norm_eps),
        x_emb=sd["x_position_embeddings.weight"], y_emb=sd["y_position_embeddings.weight"], ln_weight=sd["LayerNorm.weight"],
        ln_bias=sd["LayerNorm.bias"], lin_weight=sd["spatial_emb_matcher.layers.0.weight"],
        lin_bias=sd["spatial_emb_matcher.layers.0.bias"], **out)
    mpath = os.path.join(GOLDEN, "MANIFEST.json")
    manifest = json.load(open(mpath))
    manifest["files"]["vt5_embed.npz"] = ("SpatialEmbeddings.forward (src/_modules.py:48-86, eval) with seeded weights + the embedding "
                                          "sum of VT5.prepare_inputs_for_vqa (src/VT5.py:141-206) run as written; torch %s" % torch.__version__)
    json.dump(manifest, open(mpath, "w"), indent=1)
    print("wrote vt5_embed.npz", os.path.getsize(os.path.join(GOLDEN, "vt5_embed.npz")), "bytes; state dict keys:", sorted(sd))


if __name__ == "__main__":
    main()
