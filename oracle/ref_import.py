"""TEST INFRASTRUCTURE ONLY -- imports the *real* reference (Pikurrot/RAG-DocVQA) on CPU.

Only usable in the build container, where /root/reference exists (it does not exist on the
GPU box).  Used by oracle/make_golden.py to freeze golden vectors under tests/golden/ and by
tests that pin oracle/ref_restated.py against the reference when it is present.

The reference's src/_modules.py imports packages that are absent here (sentence_transformers,
FlagEmbedding, doclayout_yolo) and one symbol transformers 5.x dropped (render_header, used by
src/custom_pix2struct_processor.py:10-13).  None of them is touched by the retrieval path, so
they are replaced by empty stub modules before import.
"""
import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("RDV_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "src", "_modules.py"))


def _stub(name: str, **attrs):
    mod = types.ModuleType(name)
    for key, val in attrs.items():
        setattr(mod, key, val)
    sys.modules[name] = mod
    return mod


def import_reference():
    """Returns (src._modules, src.utils, src._model_utils) of the unmodified reference."""
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    sys.dont_write_bytecode = True  # the reference tree is read-only
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)

    class _Missing:  # any attempt to *use* a stubbed dependency must fail loudly
        def __init__(self, *a, **k):
            raise RuntimeError("stubbed third-party dependency used on the retrieval path")

    for name, attrs in (
        ("doclayout_yolo", dict(YOLOv10=_Missing)),
        ("sentence_transformers", dict(SentenceTransformer=_Missing, CrossEncoder=_Missing)),
        ("FlagEmbedding", dict(FlagLLMReranker=_Missing)),
    ):
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:
                _stub(name, **attrs)
    if "src.custom_pix2struct_processor" not in sys.modules:
        try:
            importlib.import_module("src.custom_pix2struct_processor")
        except Exception:
            sys.modules.pop("src.custom_pix2struct_processor", None)
            _stub("src.custom_pix2struct_processor", extract_flattened_patches_single=_Missing)
    modules = importlib.import_module("src._modules")
    utils = importlib.import_module("src.utils")
    model_utils = importlib.import_module("src._model_utils")
    return modules, utils, model_utils
