"""ctypes binding of librdv.so (include/rdv.h) -- the only door to the CUDA kernels.

There is no CPU fallback: if the library is missing, importing this module raises.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int32, c_int64, c_void_p, POINTER, Structure

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "librdv.so")
ABI_VERSION = 19

OK, E_INVALID, E_ALIGN, E_CUDA, E_LIMIT = 0, -1, -2, -3, -4
SCORE_AUTO, SCORE_LDG, SCORE_TMA = 0, 1, 2


class RdvError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__("librdv error %d: %s" % (code, message))
        self.code = code


class DocStoreStruct(Structure):
    """Mirror of `rdv_docstore` (include/rdv.h)."""
    _fields_ = [("B", c_int32), ("reserved", c_int32)] + [(n, c_void_p) for n in (
        "chunk_rec", "chunk_off", "chunk_word_off", "word_tok_off", "tok_ids", "tok_word", "word_box", "chunk_label", "chunk_page",
        "chunk_page_start", "page_chunks", "run_begin", "run_end", "doc_page_off", "page_wh", "tok_rec", "chunk_bbox")]


class GatherArgsStruct(Structure):
    """Mirror of `rdv_gather_args` (include/rdv.h)."""
    _fields_ = [("topk_idx", c_void_p), ("topk_cnt", c_void_p), ("k", c_int32), ("include_surroundings", c_int32),
                ("reorder_chunks", c_int32), ("n_sep", c_int32), ("prompt_off", c_void_p), ("prompt_ids", c_void_p),
                ("sep_ids", c_void_p), ("eos_id", c_int32), ("pad_id", c_int32), ("max_len", c_int32),
                ("max_seg", c_int32)] + [(n, c_void_p) for n in (
                    "seg_ws", "out_ids", "out_boxes", "out_mask", "out_labels", "full_len", "status", "hit_chunk",
                    "hit_page", "hit_label", "hit_nwords", "hit_bbox", "hit_rect", "sims", "topk_val")] + [
                    ("max_rows", c_int32), ("reserved", c_int32), ("emit_order", c_void_p), ("emit_cnt", c_void_p)]


class PageStoreStruct(Structure):
    """Mirror of `rdv_pagestore` (include/rdv.h)."""
    _fields_ = [("B", c_int32), ("reserved", c_int32), ("doc_page_off", c_void_p), ("page_wh", c_void_p),
                ("page_off", c_void_p), ("pixels", c_void_p)]


class VisualArgsStruct(Structure):
    """Mirror of `rdv_visual_args` (include/rdv.h)."""
    _fields_ = [("hit_page", c_void_p), ("hit_rect", c_void_p), ("hit_cnt", c_void_p), ("k", c_int32), ("out_size", c_int32),
                ("filter", c_int32), ("ksize_cap_h", c_int32), ("ksize_cap_v", c_int32), ("rows_cap", c_int32),
                ("max_page_w", c_int32), ("reserved", c_int32), ("mean", c_float * 3), ("std", c_float * 3), ("layout", c_void_p), ("coeff_h", c_void_p),
                ("coeff_v", c_void_p), ("temp", c_void_p), ("out_u8", c_void_p), ("out_px", c_void_p), ("status", c_void_p)]


class P2SArgsStruct(Structure):
    """Mirror of `rdv_p2s_args` (include/rdv.h)."""
    _fields_ = [("images", c_void_p), ("n_images", c_int32), ("n_docs", c_int32), ("max_total", c_int32), ("patch", c_int32),
                ("do_normalize", c_int32), ("max_rw", c_int32), ("max_rwh", c_int64), ("stats", c_void_p), ("temp", c_void_p),
                ("doc_total", c_void_p), ("out", c_void_p), ("mask", c_void_p), ("max_h", c_int32), ("reserved", c_int32)]


class EmbedTablesStruct(Structure):
    """Mirror of `rdv_vt5_embed_tables` (include/rdv.h)."""
    _fields_ = [("D", c_int32), ("n_pos", c_int32)] + [(n, c_void_p) for n in ("xw", "yw", "gxx", "gxy", "gyy", "c")] + [
        ("eps", c_float), ("V", c_int32), ("shared", c_void_p), ("layout", c_void_p), ("n_labels", c_int32),
        ("layout_scale", c_float)]


class SmallLayoutStruct(Structure):
    """Mirror of `rdv_small_layout` (include/rdv.h)."""
    _fields_ = [(n, c_int32) for n in ("algo", "tile_rows", "n_tiles", "max_rows")] + [(n, c_int64) for n in (
        "total_rows", "n_ctas", "cluster", "slice_rows", "o_tiles", "o_ctas", "o_q", "o_emb", "in_bytes", "o_idx", "o_cnt", "read_bytes", "o_val", "out_bytes")]


SMALL_GROW = 1
SMALL_CLUSTER = 3          # rdv_small_layout.algo: the batch takes the one-launch cluster kernel
S2_COMBINED, S2_SPATIAL, S2_SEMANTIC = 0, 1, 2

# name -> (restype, argtypes); tests/test_abi.py checks this table against include/rdv.h
SIGNATURES = {
    "rdv_abi_version": (c_int32, []),
    "rdv_last_error": (c_char_p, []),
    "rdv_device_info": (c_int32, [POINTER(c_int32), POINTER(c_int32), POINTER(c_int32)]),
    "rdv_score_plan": (c_int32, [c_int64, c_int32, c_int32, POINTER(c_int32), POINTER(c_int32)]),
    "rdv_count_tiles": (c_int64, [c_void_p, c_int32, c_int32]),
    "rdv_build_doc_table": (c_int32, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_int64,
                                      POINTER(c_int32)]),
    "rdv_upload_docs_f32": (c_int32, [c_void_p, c_void_p, c_int32, c_int32, c_void_p, c_void_p]),
    "rdv_small_batch_layout": (c_int32, [c_void_p, c_int32, c_int32, c_int32, c_void_p]),
    "rdv_small_batch_pack": (c_int32, [c_void_p, c_void_p, c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_void_p]),
    "rdv_retrieve_small_f32": (c_int32, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_void_p,
                                         c_int64, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p]),
    "rdv_score_topk_f32": (c_int32, [c_void_p, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_int32, c_int32,
                                     c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "rdv_cluster_plan": (c_int32, [c_void_p, c_int32, c_int32, c_int32, c_int32, POINTER(c_int32), POINTER(c_int32),
                                   POINTER(c_int64)]),
    "rdv_cluster_max_rows": (c_int32, [c_int32]),
    "rdv_cluster_max_k": (c_int32, []),
    "rdv_retrieve_plan": (c_int32, [c_int64, c_int32, c_int32, c_int32, c_int32, POINTER(c_int32)]),
    "rdv_debug_trace": (c_int32, [c_void_p]),
    "rdv_cluster_table_size": (c_int64, [c_void_p, c_int32, c_int32, c_int32]),
    "rdv_build_cluster_table": (c_int32, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p, c_int64]),
    "rdv_score_topk_cluster_f32": (c_int32, [c_void_p, c_int64, c_int32, c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p,
                                             c_void_p, c_void_p, c_void_p, c_void_p]),
    "rdv_retrieve_vt5_f32": (c_int32, [c_void_p, c_int64, c_int32, c_void_p, c_int32, c_int32, c_void_p,
                                       POINTER(DocStoreStruct), POINTER(GatherArgsStruct), c_void_p]),
    "rdv_score_f32": (c_int32, [c_void_p, c_int32, c_int32, c_int32, c_void_p, c_int32, c_int32, c_void_p, c_void_p]),
    "rdv_topk_segments_f32": (c_int32, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_void_p, c_void_p,
                                        c_void_p, c_void_p]),
    "rdv_pooled_select_f32": (c_int32, [c_void_p, c_void_p, c_int64, c_int32, c_void_p, c_int32, c_int32, c_int32, c_int32,
                                        c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                        c_void_p, c_void_p, c_void_p]),
    "rdv_struct_size": (c_int64, [c_char_p]),
    "rdv_vt5_embed_tables_build": (c_int32, [c_void_p, c_void_p, c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_void_p,
                                             c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "rdv_vt5_input_embeds_f32": (c_int32, [POINTER(EmbedTablesStruct), c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int64,
                                           c_void_p, c_int64, c_void_p, c_void_p, c_void_p]),
    "rdv_mean_pool_f32": (c_int32, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p, c_void_p,
                                    c_void_p, c_void_p]),
    "rdv_row_inv_norm_f32": (c_int32, [c_void_p, c_int64, c_int32, c_void_p, c_void_p]),
    "rdv_maxsim_f32": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32,
                                 c_void_p, c_void_p, c_void_p, c_void_p]),
    "rdv_maxsim_tiles_i": (c_int32, [c_int32]),
    "rdv_topk_merge": (c_int32, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_void_p]),
    "rdv_topk_merge_parts": (c_int32, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int64, c_int64, c_int32, c_void_p,
                                       c_void_p, c_void_p]),
    "rdv_rows_to_bf16": (c_int32, [c_void_p, c_int64, c_int32, c_int32, c_void_p, c_void_p, c_void_p]),
    "rdv_bf16_inv_norm": (c_int32, [c_void_p, c_int64, c_int32, c_void_p, c_void_p]),
    "rdv_corpus_groups": (c_int32, [c_int64, c_int32]),
    "rdv_tc_tile_m": (c_int32, []),
    "rdv_tc_candidates_per_group": (c_int32, []),
    "rdv_corpus_score_topk_bf16": (c_int32, [c_void_p, c_void_p, c_int64, c_int32, c_void_p, c_void_p, c_int32, c_int32,
                                             c_int64, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "rdv_maxsim_bf16_tc": (c_int32, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p, c_void_p,
                                     c_void_p]),
    "rdv_rows_split_tf32": (c_int32, [c_void_p, c_int64, c_int32, c_int32, c_void_p, c_void_p, c_void_p]),
    "rdv_maxsim_tf32x3_tc": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32,
                                       c_void_p, c_void_p, c_void_p]),
    "rdv_visual_pack": (c_int32, [POINTER(PageStoreStruct), POINTER(VisualArgsStruct), c_void_p]),
    "rdv_pix2struct_patches": (c_int32, [POINTER(PageStoreStruct), POINTER(P2SArgsStruct), c_void_p]),
    "rdv_rerank_order": (c_int32, [c_void_p, c_int32, c_void_p, c_int32, c_int32, ctypes.c_double, c_int32, c_int32,
                                   c_void_p, c_void_p, c_void_p, c_void_p]),
    "rdv_page_vote": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p,
                                c_void_p, c_void_p]),
    "rdv_layout_assign": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32,
                                    c_void_p, c_void_p, c_void_p, c_void_p]),
    "rdv_s2_weights": (c_int32, [c_void_p, c_void_p, c_int32, c_void_p, c_int32, c_int32, c_void_p, c_int64, c_void_p,
                                 c_void_p]),
    "rdv_gather_vt5_inputs": (c_int32, [POINTER(DocStoreStruct), POINTER(GatherArgsStruct), c_void_p]),
}


def _load() -> ctypes.CDLL:
    if not os.path.exists(LIB_PATH):
        # a fresh checkout: compile the CUDA sources once (nvcc); without a compiler there is nothing to fall back to
        try:
            from . import build as _build
            _build.build()
        except Exception as exc:
            raise ImportError(
                "%s not found and could not be built (%s): run `python -m rag_docvqa_b200.build` (needs nvcc). "
                "rag_docvqa_b200 has no CPU fallback." % (LIB_PATH, exc))
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
