"""The drop-in boundary on the host side (no GPU): signatures equal the reference's, and
compat.install() rebinds the reference's names when the reference tree is importable."""
import inspect

import pytest

from oracle.ref_import import reference_available


def params(fn):
    return [p for p in inspect.signature(fn).parameters if p != "self"]


def test_signatures_match_reference_call_sites():
    # reference: src/_modules.py:2155-2164 (Retriever.retrieve), :2453-2461 (VisualRetriever.retrieve),
    # src/_model_utils.py:49-52 (mean_pooling), src/utils.py:442 (late_interaction)
    from rag_docvqa_b200 import functional as F
    from rag_docvqa_b200.retriever import Retriever, VisualRetriever
    assert params(Retriever.retrieve) == ["text_embeddings", "question_embeddings", "words_text_chunks",
                                          "words_box_chunks", "layout_labels_chunks", "images", "page_indices"]
    assert params(Retriever._get_similarities) == ["text_embeddings", "question_embeddings"]
    assert params(Retriever._get_top_k) == ["similarities", "words_text_chunks", "words_box_chunks",
                                            "layout_labels_chunks", "images", "page_indices"]
    assert params(VisualRetriever.retrieve) == ["patch_embeddings", "question_embeddings", "patches_flatten_indices",
                                                "patches_matrix_list", "patches_xyxy", "images"]
    assert params(F.mean_pooling)[:2] == ["embs", "attention_mask"]
    assert params(F.late_interaction)[:2] == ["query", "patches"]      # + optional `mode` (default keeps the contract)
    assert all(p.default is not inspect.Parameter.empty
               for p in list(inspect.signature(F.late_interaction).parameters.values())[2:])


@pytest.mark.skipif(not reference_available(), reason="reference tree not mounted (GPU box)")
def test_signatures_equal_live_reference_and_install_rebinds():
    from oracle.ref_import import import_reference
    modules, utils, model_utils = import_reference()
    from rag_docvqa_b200 import compat
    from rag_docvqa_b200.retriever import Retriever, VisualRetriever
    assert params(Retriever.retrieve) == params(modules.Retriever.retrieve)
    assert params(Retriever.__init__) == params(modules.Retriever.__init__)
    assert params(Retriever._get_top_k) == params(modules.Retriever._get_top_k)
    assert params(VisualRetriever.retrieve) == params(modules.VisualRetriever.retrieve)
    assert params(VisualRetriever._get_top_k) == params(modules.VisualRetriever._get_top_k)
    from rag_docvqa_b200.postproc import Reranker
    assert params(Reranker.rerank) == params(modules.Reranker.rerank)                  # src/_modules.py:1558-1563
    assert params(Reranker.batch_rerank) == params(modules.Reranker.batch_rerank)      # :1597-1602
    assert params(Reranker.__init__) == params(modules.Reranker.__init__)              # :1541-1545
    from rag_docvqa_b200.chunker import Chunker
    assert params(Chunker.get_chunks) == params(modules.Chunker.get_chunks)            # :872-878
    assert params(Chunker.__init__) == params(modules.Chunker.__init__)
    from rag_docvqa_b200.s2chunker import S2Chunker
    for name in ("__init__", "forward", "create_nodes_and_edges", "cluster", "_combined_weights", "_spatial_weights_calculation",
                 "_semantic_weights_calculation", "_calculate_n_clusters", "_cluster_graph", "_split_clusters_by_token_length"):
        assert params(getattr(S2Chunker, name)) == params(getattr(modules.S2Chunker, name)), name    # :1669-1962
    ref_retriever = modules.Retriever
    try:
        done = compat.install()
        assert ("src._modules", "Retriever") in done and ("src._modules", "late_interaction") in done
        assert modules.Retriever is Retriever and modules.VisualRetriever is VisualRetriever
        assert ("src._modules", "S2Chunker") in done and issubclass(modules.S2Chunker, S2Chunker)
        assert utils.late_interaction.__doc__.startswith("reference signature")
        assert compat.install() == []          # idempotent
    finally:
        compat.uninstall()
    assert modules.Retriever is ref_retriever


def test_cpu_tensors_are_rejected_not_silently_computed():
    import torch
    from rag_docvqa_b200 import functional as F
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        F.mean_pooling(torch.zeros(2, 3, 4), torch.ones(2, 3, dtype=torch.long))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        F.late_interaction(torch.zeros(1, 3, 4), torch.zeros(2, 3, 4))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        F.topk_merge(torch.zeros(2, 4), torch.zeros(2, 4, dtype=torch.long), 2)


def test_lazy_crop_uses_public_pillow_protocol_only():
    """The deferred patch must behave like page.crop(rect) with the reference's pinned Pillow 10.3 as well as with the
    installed one: no private Pillow field (`_im` exists only from Pillow 11) is read or written."""
    import inspect

    import numpy as np
    from PIL import Image

    from rag_docvqa_b200 import retriever
    src = inspect.getsource(retriever._lazy_crop_class)
    code = "\n".join(line.split("#")[0] for line in src.splitlines())
    assert "_im" not in code.replace("_image", "")
    rng = np.random.RandomState(0)
    page = Image.fromarray(rng.randint(0, 256, (60, 80, 3)).astype(np.uint8), "RGB")
    page.info["dpi"] = (72, 72)
    for rect in ((5, 7, 40, 33), (0, 0, 80, 60), (-3, -2, 10, 12), (70, 50, 95, 70)):
        lazy = retriever.lazy_crop(page, rect)
        want = page.crop(rect)
        assert lazy.size == want.size and lazy.mode == want.mode and lazy.info == want.info
        assert lazy._page is not None                                   # nothing cut yet
        assert np.array_equal(np.asarray(lazy), np.asarray(want))        # first pixel access cuts
        assert lazy._page is None
        assert lazy.resize((8, 8)).size == (8, 8) and lazy.copy().tobytes() == want.tobytes()
    pal = page.convert("P")
    assert retriever.lazy_crop(pal, (1, 1, 9, 9)).convert("RGB").tobytes() == pal.crop((1, 1, 9, 9)).convert("RGB").tobytes()


def test_gc_pause_is_opt_in(monkeypatch):
    """Retriever.retrieve leaves the process-wide collector alone unless `retrieval_pause_gc` asks for the pause."""
    import gc

    from rag_docvqa_b200.retriever import Retriever
    seen = []
    base = {"compute_stats": False, "compute_stats_examples": False, "n_stats_examples": 0, "device": "cuda:0"}
    monkeypatch.setattr(Retriever, "_retrieve", lambda self, *a: seen.append(gc.isenabled()) or "out")
    assert gc.isenabled()
    assert Retriever(base).retrieve(None, None, None, None, None, None, None) == "out"
    assert Retriever({**base, "retrieval_pause_gc": True}).retrieve(None, None, None, None, None, None, None) == "out"
    assert seen == [True, False] and gc.isenabled()
