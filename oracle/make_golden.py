"""TEST INFRASTRUCTURE ONLY -- freezes outputs of the UNMODIFIED reference as golden vectors.

Run in the build container (where /root/reference exists):

    PYTHONDONTWRITEBYTECODE=1 python -m oracle.make_golden

Writes tests/golden/*.npz / *.json.  Inputs come from rag_docvqa_b200/synth.py with fixed seeds and
are stored alongside the outputs, so the GPU box (no reference tree) replays them byte-for-byte.
torch used to generate: recorded in tests/golden/MANIFEST.json.
"""
from __future__ import annotations

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
from rag_docvqa_b200 import synth  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
BASE_CFG = {"compute_stats": False, "compute_stats_examples": False, "n_stats_examples": 0}


def image_digest(im):
    return [im.size[0], im.size[1], zlib.crc32(im.convert("RGB").tobytes()) & 0xFFFFFFFF]


def text_case(name, sizes, dim, k, seed, normalised, chunks_per_page):
    emb, q = synth.make_embeddings(sizes, dim, seed, normalised=normalised, dup_frac=0.05)
    return dict(name=name, sizes=sizes, dim=dim, k=k, seed=seed, emb=emb, q=q,
                pages=synth.make_page_indices(sizes, chunks_per_page))


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    modules, utils, model_utils = import_reference()
    torch.manual_seed(0)
    manifest = {"torch": torch.__version__, "numpy": np.__version__, "files": {}}

    # ---- a4 + a6: score and torch.topk on C1 and on a ragged batch --------------------------------
    cases = [
        text_case("c1", [30], 384, 5, synth.SEED_BASE + 1, True, 30),
        text_case("ragged_norm", [150, 60, 7, 0, 90, 3, 121, 1], 384, 5, 4242, True, 30),
        text_case("ragged_raw", [64, 0, 33, 200, 2, 10], 768, 10, 4243, False, 50),
        text_case("k20_d1024", [77, 19, 140], 1024, 20, 4244, True, 25),
    ]
    for case in cases:
        retr = modules.Retriever({**BASE_CFG, "chunk_num": case["k"]})
        sims = retr._get_similarities(case["emb"], case["q"])
        topk = [torch.topk(s, k=min(case["k"], len(s))).indices.numpy() for s in sims]
        arrays = {"sizes": np.array(case["sizes"], dtype=np.int64), "k": np.array(case["k"]),
                  "q": case["q"].numpy()}
        for b, (e, s, t) in enumerate(zip(case["emb"], sims, topk)):
            arrays["emb_%d" % b] = e.numpy()
            arrays["sims_%d" % b] = s.numpy()
            arrays["topk_%d" % b] = t
        path = "score_topk_%s.npz" % case["name"]
        np.savez_compressed(os.path.join(GOLDEN, path), **arrays)
        manifest["files"][path] = "Retriever._get_similarities + torch.topk (src/_modules.py:1978-1997, 2015-2016)"

    # ---- a7/a8/a9/a11: full Retriever.retrieve list outputs --------------------------------------
    sizes = [24, 9, 0, 3, 40]
    cpp = 8
    emb, q = synth.make_embeddings(sizes, 64, 777, normalised=True, dup_frac=0.0)
    words, boxes, labels = synth.make_words(sizes, 778, min_words=3, max_words=9, empty_chunk_every=11)
    pages = synth.make_page_indices(sizes, cpp)
    images = synth.make_images(sizes, cpp, width=170, height=220, ragged_sizes=True)
    np.savez_compressed(os.path.join(GOLDEN, "retrieve_inputs.npz"), q=q.numpy(),
                        **{"emb_%d" % b: e.numpy() for b, e in enumerate(emb)})
    retrieve_out = {"sizes": sizes, "chunks_per_page": cpp, "dim": 64, "emb_seed": 777, "words_seed": 778,
                    "image_wh": [170, 220], "variants": []}
    for k in (4, 6):
        for s in (0, 3):
            for reorder in (False, True):
                retr = modules.Retriever({**BASE_CFG, "chunk_num": k, "include_surroundings": s,
                                          "reorder_chunks": reorder})
                out = retr.retrieve(emb, q, words, boxes, labels, images, pages)
                sims = out[8]
                retrieve_out["variants"].append({
                    "k": k, "include_surroundings": s, "reorder_chunks": reorder,
                    "topk": [torch.topk(x, k=min(k, len(x))).indices.tolist() for x in sims],
                    "top_k_text": out[0], "top_k_boxes": out[1], "top_k_layout_labels": out[2],
                    "top_k_words_text": out[3], "top_k_words_boxes": out[4],
                    "top_k_words_layout_labels": out[5],
                    "top_k_patches": [[image_digest(im) for im in doc] for doc in out[6]],
                    "top_k_page_indices": out[7],
                    "similarities": [x.tolist() for x in sims],
                })
    with open(os.path.join(GOLDEN, "retrieve_lists.json"), "w") as f:
        json.dump(retrieve_out, f)
    manifest["files"]["retrieve_lists.json"] = "Retriever.retrieve 9-tuple (src/_modules.py:2155-2180)"
    manifest["files"]["retrieve_inputs.npz"] = "embeddings for retrieve_lists.json"

    # ---- a1: mean_pooling ------------------------------------------------------------------------
    embs, mask = synth.make_token_batch(9, 40, 99, mean_len=12, std_len=5, min_len=1, max_len=24, all_pad_rows=2)
    pooled = model_utils.mean_pooling(embs, mask)
    np.savez_compressed(os.path.join(GOLDEN, "mean_pooling.npz"), embs=embs.numpy(), mask=mask.numpy(),
                        pooled=pooled.numpy())
    manifest["files"]["mean_pooling.npz"] = "mean_pooling (src/_model_utils.py:49-61)"

    # ---- a5: late_interaction --------------------------------------------------------------------
    patches, qv = synth.make_strip_batch(2, [5, 3], 48, 96, 314)
    patches[1][1, 7] = 0.0   # an all-zero strip token (F.normalize eps path)
    vr = modules.VisualRetriever({"chunk_num": 2})
    vs = vr._get_similarities(patches, qv)
    np.savez_compressed(os.path.join(GOLDEN, "late_interaction.npz"), q=qv.numpy(),
                        p0=patches[0].numpy(), p1=patches[1].numpy(), s0=vs[0].numpy(), s1=vs[1].numpy())
    manifest["files"]["late_interaction.npz"] = "late_interaction via VisualRetriever._get_similarities (src/utils.py:442-458)"

    # ---- a10: VisualRetriever.retrieve decode ----------------------------------------------------
    from PIL import Image
    rng = np.random.RandomState(5)
    vis = {"docs": []}
    strips_per_group = [[4, 3, 5], [2], []]
    patch_h, overlap = 40, 8
    v_patches, v_q, v_flat, v_mats, v_xyxy, v_images = [], [], [], [], [], []
    for b, groups in enumerate(strips_per_group):
        flat, mats, xyxy, imgs = [], [], [], []
        for g, n_rows in enumerate(groups):
            W, H = 120 + 10 * g, (patch_h - overlap) * n_rows + overlap
            arr = rng.randint(0, 255, size=(H, W, 3)).astype(np.uint8)
            page = Image.fromarray(arr, "RGB")
            rows = [[0, r * (patch_h - overlap), W, r * (patch_h - overlap) + patch_h] for r in range(n_rows)]
            mats.append([[page.crop(tuple(rc))] for rc in rows])
            xyxy.append(rows)
            flat.extend([g] * n_rows)
            imgs.append(page)
        n = len(flat)
        v_flat.append(np.array(flat, dtype=np.int64))
        v_mats.append(mats)
        v_xyxy.append(xyxy)
        v_images.append(imgs)
        g_t = torch.Generator().manual_seed(100 + b)
        v_patches.append(torch.randn(max(n, 1), 12, 16, generator=g_t) if n else torch.zeros(1, 12, 16))
        v_q.append(torch.randn(12, 16, generator=g_t))
        vis["docs"].append({"groups": groups, "xyxy": xyxy, "image_wh": [[im.size[0], im.size[1]] for im in imgs]})
    v_q = torch.stack(v_q)
    np.savez_compressed(os.path.join(GOLDEN, "visual_inputs.npz"), q=v_q.numpy(),
                        **{"p_%d" % b: p.numpy() for b, p in enumerate(v_patches)})
    vis["variants"] = []
    for k in (1, 3):
        for s in (0, 1, 2, 3, (0, 1), (1, 2)):
            vr = modules.VisualRetriever({"chunk_num": k, "include_surroundings": s, "chunk_mode": "horizontal"})
            sims = vr._get_similarities(v_patches, v_q)
            crops, page_ids = vr.retrieve(v_patches, v_q, v_flat, v_mats, v_xyxy, v_images)
            vis["variants"].append({
                "k": k, "include_surroundings": list(s) if isinstance(s, tuple) else s,
                "sims": [x.tolist() for x in sims],
                "crops": [sorted(image_digest(im) for im in doc) for doc in crops],
                "pages": [sorted(doc) for doc in page_ids],
            })
    with open(os.path.join(GOLDEN, "visual_retrieve.json"), "w") as f:
        json.dump(vis, f)
    manifest["files"]["visual_retrieve.json"] = "VisualRetriever.retrieve (src/_modules.py:2386-2464)"
    manifest["files"]["visual_inputs.npz"] = "strip embeddings for visual_retrieve.json"

    # ---- a12: VT5.prepare_inputs_for_vqa (ids / boxes / mask / layout labels) -------------------
    import importlib
    vt5 = importlib.import_module("src.VT5")
    table = synth.make_tokens_for_words(words, seed=3)

    class FakeTokenizer:  # stand-in for T5Tokenizer: no tokenizer files exist offline
        eos_token_id, pad_token_id = 1, 0

        def __call__(self, text, **kw):
            if text.startswith("question: "):
                ids = [5 + (zlib.crc32(t.encode()) % 1000) for t in text.split()]
            else:
                ids = list(table.get(text, [2]))
            return types.SimpleNamespace(input_ids=ids + [self.eos_token_id])

    packs = {"variants": []}
    retr = modules.Retriever({**BASE_CFG, "chunk_num": 6})
    out = retr.retrieve(emb, q, words, boxes, labels, images, pages)
    questions = ["what is item %d about ?" % b for b in range(len(sizes))]
    for use_layout, max_len, sep in (("Default", 512, None), ("Embed", 512, None), ("Default", 48, None),
                                     ("Embed", 40, "<sep>")):
        fake_self = types.SimpleNamespace(
            tokenizer=FakeTokenizer(), max_source_length=max_len, use_layout_labels=use_layout,
            language_backbone=types.SimpleNamespace(device="cpu", shared=lambda ids: torch.zeros(*ids.shape, 1)),
            spatial_embedding=lambda bx: torch.zeros(*bx.shape[:2], 1),
            layout_embedding=lambda lb: torch.zeros(*lb.shape, 1), layout_embedding_scale=1.0,
            visual_embedding=lambda ims: (torch.zeros(len(ims), 0, 1), torch.zeros(len(ims), 0, dtype=torch.long)))
        w_flat = [utils.flatten(b, sep) for b in out[3]]
        b_flat = [utils.flatten(b, sep) for b in out[4]]
        l_flat = [utils.flatten(b, sep) for b in out[5]]
        captured = {}
        real_spatial = fake_self.spatial_embedding

        def spy(bx, _c=captured, _r=real_spatial):
            _c["boxes"] = bx.clone()
            return _r(bx)
        fake_self.spatial_embedding = spy
        res = vt5.VT5ForConditionalGeneration.prepare_inputs_for_vqa(
            fake_self, questions, w_flat, b_flat, l_flat, [None] * len(sizes), None, return_ids=True)
        input_ids, attn, _labels, layout = res
        packs["variants"].append({
            "use_layout_labels": use_layout, "max_source_length": max_len, "sep": sep,
            "input_ids": input_ids.tolist(), "attention_mask": attn.tolist(),
            "boxes": captured["boxes"].tolist(), "layout_labels": None if layout is None else layout.tolist()})
    packs["questions"] = questions
    packs["k"] = 6
    packs["word_table_seed"] = 3
    with open(os.path.join(GOLDEN, "vt5_pack.json"), "w") as f:
        json.dump(packs, f)
    manifest["files"]["vt5_pack.json"] = "flatten + VT5.prepare_inputs_for_vqa ids/boxes/mask (src/utils.py:233-253, src/VT5.py:141-192)"

    with open(os.path.join(GOLDEN, "MANIFEST.json"), "w") as f:
        json.dump(manifest, f, indent=1)
    for name in sorted(os.listdir(GOLDEN)):
        print("%9d  %s" % (os.path.getsize(os.path.join(GOLDEN, name)), name))


if __name__ == "__main__":
    main()
