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
