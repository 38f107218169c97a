"""B200-native retrieval stage of RAG-DocVQA.

    from rag_docvqa_b200 import Retriever, VisualRetriever, mean_pooling, late_interaction      # the drop-ins
    from rag_docvqa_b200 import Reranker, major_page_indices, Chunker, S2Chunker                 # either side of the path
    from rag_docvqa_b200 import DocStore, PageStore, CorpusShard                                 # device-resident stores

Names are resolved on first use, so `import rag_docvqa_b200` itself needs neither torch nor the built library
(rag_docvqa_b200.build can be run from a bare checkout).
"""
_EXPORTS = {
    "Retriever": "retriever", "VisualRetriever": "retriever",
    "mean_pooling": "functional", "late_interaction": "functional", "score_topk": "functional",
    "Reranker": "postproc", "major_page_indices": "postproc", "Chunker": "chunker", "S2Chunker": "s2chunker",
    "DocStore": "docstore", "PageStore": "pagestore", "CorpusShard": "sharded", "CorpusIndex": "sharded",
}

__all__ = sorted(_EXPORTS)


def __getattr__(name):
    mod = _EXPORTS.get(name)
    if mod is None:
        raise AttributeError("module %r has no attribute %r" % (__name__, name))
    import importlib
    return getattr(importlib.import_module("." + mod, __name__), name)
