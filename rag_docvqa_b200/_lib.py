"""ctypes binding of librdv.so (include/rdv.h) -- the only door to the CUDA kernels.

There is no CPU fallback: if the library is missing, importing this module raises.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int32, c_int64, c_void_p, c_size_t, POINTER

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "librdv.so")
ABI_VERSION = 1

OK, E_INVALID, E_ALIGN, E_CUDA, E_LIMIT = 0, -1, -2, -3, -4


class RdvError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__("librdv error %d: %s" % (code, message))
        self.code = code


# name -> (restype, argtypes); tests/test_abi.py checks this table against include/rdv.h
SIGNATURES = {
    "rdv_abi_version": (c_int32, []),
    "rdv_last_error": (c_char_p, []),
    "rdv_device_info": (c_int32, [POINTER(c_int32), POINTER(c_int32), POINTER(c_int32)]),
    "rdv_score_topk_f32": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32,
                                     c_int32, c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_void_p,
                                     c_void_p, c_void_p]),
    "rdv_score_tile_rows": (c_int32, [c_int64, c_int32]),
}


def _load() -> ctypes.CDLL:
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "%s not found: build it with `python -m rag_docvqa_b200.build` (needs nvcc). "
            "rag_docvqa_b200 has no CPU fallback." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)      # AttributeError if the .so is stale
        fn.restype = restype
        fn.argtypes = argtypes
    got = lib.rdv_abi_version()
    if got != ABI_VERSION:
        raise ImportError("librdv.so ABI %d != binding ABI %d: rebuild with `python -m rag_docvqa_b200.build --force`"
                          % (got, ABI_VERSION))
    return lib


lib = _load()


def check(code: int) -> None:
    if code != OK:
        raise RdvError(code, (lib.rdv_last_error() or b"").decode("utf-8", "replace"))


def device_info():
    sm, major, minor = c_int32(), c_int32(), c_int32()
    check(lib.rdv_device_info(ctypes.byref(sm), ctypes.byref(major), ctypes.byref(minor)))
    return sm.value, major.value, minor.value
