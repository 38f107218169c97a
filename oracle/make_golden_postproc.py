"""TEST INFRASTRUCTURE ONLY -- freezes the reference's reranker post-processing and page vote.

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden_postproc.py

Runs, in the build container (needs /root/reference), the UNMODIFIED
  * src._modules.Reranker.rerank / batch_rerank (src/_modules.py:1562-1610) with a stand-in cross-encoder that returns
    prepared scores (the cross-encoder is a model and out of scope), and
  * src.RAGVT5.RAGVT5.forward (src/RAGVT5.py:318-520) on a stand-in `self` whose online_retrieve() returns prepared
    retrieval outputs and whose generator returns a fixed 4-tuple -- so the majorpage / weightmajorpage vote
    (:455-477) runs exactly as written,
and writes tests/golden/postproc.json.  The page vote here runs under THIS container's numpy (>= 2, NEP 50:
float32 accumulation); the reference's pinned numpy 1.26.4 accumulates in float64 -- oracle/ref_restated.page_vote
states both, the frozen values pin the NEP-50 one.
"""
import base64
import json
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.ref_import import import_reference  # noqa: E402


def main():
    modules, _, _ = import_reference()
    # src/RAGVT5.py:20 imports the Qwen generator wrapper, whose dependencies (qwen_vl_utils, peft) are absent here and
    # which the page vote never touches: replaced by a stub module, like the other absent third-party packages
    if "src.QwenVLInstruct" not in sys.modules:
        stub = types.ModuleType("src.QwenVLInstruct")
        stub.QwenVLForConditionalGeneration = None
        sys.modules["src.QwenVLInstruct"] = stub
    import importlib
    ragvt5 = importlib.import_module("src.RAGVT5")
    rng = np.random.RandomState(20261018)

    # ---- Reranker.rerank ---------------------------------------------------------------------------------
    class FixedScores:
        def __init__(self):
            self.next = None

        def forward(self, pairs):
            assert len(pairs) == len(self.next)
            return self.next

    rerank_cases = []
    settings = [(0.4, 5, 1), (0.4, 3, 1), (0.4, 5, 2), (0.9, 5, 3), (0.0, 10, 1), (0.5, 2, 4), (0.4, 5, 0)]
    for case in range(60):
        k = int(rng.choice([1, 2, 3, 5, 8, 10, 16]))     # <= 16: numpy's argsort is an insertion sort (stable)
        kind = case % 4
        if kind == 0:
            s = rng.rand(k).astype(np.float32)
        elif kind == 1:                                   # ties
            s = rng.choice(np.array([0.1, 0.4, 0.55, 0.9], dtype=np.float32), size=k)
        elif kind == 2:                                   # logits (FlagLLMReranker returns a list of floats)
            s = [float(x) for x in rng.randn(k) * 3]
        else:                                             # torch tensor (the empty-input branch returns one; :1510)
            s = torch.from_numpy(rng.rand(k).astype(np.float32))
        thresh, mx, mn = settings[case % len(settings)]
        ce = FixedScores()
        ce.next = s
        rr = modules.Reranker({"rerank_filter_tresh": thresh, "rerank_max_chunk_num": mx, "rerank_min_chunk_num": mn},
                              cross_encoder=ce)
        cands = ["c%d" % i for i in range(k)]
        ids = list(range(k))
        out_c, out_ids = rr.rerank("q", cands, ids)
        assert out_c == ["c%d" % i for i in out_ids]
        rerank_cases.append({
            "scores": [float(x) for x in (s.tolist() if hasattr(s, "tolist") else s)],
            "dtype": "f64" if isinstance(s, list) else "f32",
            "thresh": thresh, "max": mx, "min": mn, "order": [int(i) for i in out_ids]})

    # ---- page vote through RAGVT5.forward --------------------------------------------------------------------
    class Generator:
        def __call__(self, new_batch, return_pred_answer=True):
            self.seen = new_batch
            bs = len(new_batch["questions"])
            return (None, ["a"] * bs, None, [1.0] * bs)

    vote_cases = []
    for mode in ("majorpage", "weightmajorpage"):
        for case in range(30):
            bs = 4
            n_pages = int(rng.choice([3, 9, 40, 200]))
            hits, sims, pages = [], [], []
            for b in range(bs):
                n_b = int(rng.choice([0, 2, 7, 30, 600])) if case % 5 == 0 else int(rng.randint(5, 700))
                k_b = min(int(rng.choice([5, 10, 20])), n_b)
                if case % 3 == 0:      # few distinct pages: ties between pages are common in majorpage
                    pg = rng.choice(rng.randint(0, n_pages, size=3), size=k_b).tolist()
                else:
                    pg = rng.randint(0, n_pages, size=k_b).tolist()
                s = (rng.rand(n_b).astype(np.float32) * 0.8 + 0.1)
                if case % 4 == 1 and n_b:
                    s[rng.randint(0, n_b, size=max(1, n_b // 3))] *= -1     # negative cosines exist
                pages.append([int(p) for p in pg]); sims.append(s); hits.append(k_b)
            gen = Generator()
            me = types.SimpleNamespace(
                use_RAG=True, use_layout_labels="Default", add_sep_token=False, page_retrieval=mode, train_mode=False,
                train_generator=False, generator=gen, use_not_answerable_classifier=False, model_path="vt5")
            me.online_retrieve = lambda batch, return_steps=False: (
                [["t"] * h for h in hits], [[[0, 0, 1, 1]] * h for h in hits], [[1] * h for h in hits],
                [[None] * h for h in hits], [list(p) for p in pages], [[["w"]] * h for h in hits],
                [[[[0, 0, 1, 1]]] * h for h in hits], [[[1]] * h for h in hits],
                [[[1]] * n_pages for _ in range(bs)], [torch.from_numpy(s) for s in sims], {},
                {"stats": {}, "stats_examples": {}})
            batch = {"questions": ["q"] * bs, "answers": [["a"]] * bs,
                     "words": [[["w"]] * n_pages for _ in range(bs)], "boxes": [[[[0, 0, 1, 1]]] * n_pages for _ in range(bs)],
                     "images": [[None] * n_pages for _ in range(bs)]}
            import warnings
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                out = ragvt5.RAGVT5.forward(me, batch, return_pred_answer=True, return_retrieval=True)
            major = out[-1]["page_indices"]
            vote_cases.append({"mode": mode, "pages": pages, "sims_f32_b64": [base64.b64encode(s.tobytes()).decode() for s in sims],
                               "major": [int(p) for p in major], "numpy": np.__version__})

    path = os.path.join(ROOT, "tests", "golden", "postproc.json")
    with open(path, "w") as f:
        json.dump({"rerank": rerank_cases, "page_vote": vote_cases}, f)
    print("wrote", path, os.path.getsize(path), "bytes;", len(rerank_cases), "rerank cases,", len(vote_cases), "vote cases")


if __name__ == "__main__":
    main()
