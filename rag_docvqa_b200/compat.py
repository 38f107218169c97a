"""Swaps the B200 path into an importable reference tree (Pikurrot/RAG-DocVQA) without editing it.

    import rag_docvqa_b200.compat as compat
    compat.install()          # before or after `import src.RAGVT5` / `import src.RAGPix2Struct`

After install(), src/RAGVT5.py:105 (`self.retriever = Retriever(config)`) and src/RAGPix2Struct.py:83
(`VisualRetriever(config)`) construct the drop-ins of rag_docvqa_b200.retriever, and src._modules'
`mean_pooling` / `late_interaction` names point at the CUDA implementations.  The permanent form of the
same change is the four-line patch in INTEGRATION.md.
"""
from __future__ import annotations

import sys

_PATCHED = {}
_REPLACEMENTS = None


def replacements():
    global _REPLACEMENTS
    if _REPLACEMENTS is None:
        _REPLACEMENTS = _make_replacements()
    return _REPLACEMENTS


def _make_replacements():
    from . import functional, retriever
    return {
        "Retriever": retriever.Retriever,
        "VisualRetriever": retriever.VisualRetriever,
        "mean_pooling": _mean_pooling_dispatch(functional),
        "late_interaction": _late_interaction_dispatch(functional),
        "Reranker": _reranker_class(),
        "Chunker": _chunker_class(),
        "S2Chunker": _s2chunker_class(),
    }


def _s2chunker_class():
    from . import s2chunker

    class S2Chunker(s2chunker.S2Chunker):
        """reference constructor (src/_modules.py:1670-1685): in "spatial+semantic" mode without an embedder the
        REFERENCE's own BiEncoder(config) is built (the encoder stays the reference's)."""

        def __init__(self, config: dict, embedder=None):
            if embedder is None and config.get("cluster_mode", "spatial+semantic") == "spatial+semantic":
                ref = sys.modules.get("src._modules")
                if ref is None:
                    raise ValueError("S2Chunker: no embedder given and src._modules is not imported")
                embedder = ref.BiEncoder(config)
            super().__init__(config, embedder)

    return S2Chunker


def _chunker_class():
    from .chunker import Chunker
    return Chunker


def _reranker_class():
    from . import postproc

    class Reranker(postproc.Reranker):
        """reference constructor (src/_modules.py:1541-1556): without a cross-encoder argument the REFERENCE's own
        CrossEncoder / FlagLLMReranker model class is built (the models stay the reference's); everything after its
        forward() runs through rdv_rerank_order."""

        def __init__(self, config: dict, cross_encoder=None):
            if cross_encoder is None:
                ref = sys.modules.get("src._modules")
                if ref is None:
                    raise ValueError("Reranker: no cross-encoder given and src._modules is not imported")
                cls = ref.FlagLLMReranker if "gemma" in config.get("reranker_weights", "") else ref.CrossEncoder
                cross_encoder = cls(config)
            super().__init__(config, cross_encoder)

    return Reranker


def _mean_pooling_dispatch(functional):
    def mean_pooling(embs, attention_mask):
        """reference signature (src/_model_utils.py:49); CUDA tensors only -- no CPU fallback."""
        return functional.mean_pooling(embs, attention_mask)
    return mean_pooling


def _late_interaction_dispatch(functional):
    def late_interaction(query, patches):
        """reference signature (src/utils.py:442); CUDA tensors only -- no CPU fallback."""
        return functional.late_interaction(query, patches)
    return late_interaction


def install(modules=None) -> list:
    """Rebinds the seven names in every loaded `src.*` module of the reference that defines or imported
    them.  Returns the list of (module, name) pairs patched.  Idempotent; undo with uninstall()."""
    repl = replacements()
    done = []
    for mod_name, mod in list(sys.modules.items()):
        if mod is None or not (mod_name == "src" or mod_name.startswith("src.")):
            continue
        if modules is not None and mod_name not in modules:
            continue
        for name, new in repl.items():
            if hasattr(mod, name) and getattr(mod, name) is not new:
                _PATCHED.setdefault((mod_name, name), getattr(mod, name))
                setattr(mod, name, new)
                done.append((mod_name, name))
    return done


def uninstall() -> None:
    for (mod_name, name), old in list(_PATCHED.items()):
        mod = sys.modules.get(mod_name)
        if mod is not None:
            setattr(mod, name, old)
        del _PATCHED[(mod_name, name)]
