"""CPU-side checks of the C-ABI boundary: librdv.so builds, loads, and exports exactly the
symbols include/rdv.h declares (no compute is launched here)."""
import ctypes
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "rdv.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"RDV_API\s+[\w\s\*]+?\b(rdv_\w+)\s*\(", text)))


def test_header_declares_something():
    syms = declared_symbols()
    assert "rdv_score_topk_f32" in syms and "rdv_last_error" in syms


def test_library_builds_and_exports_every_declared_symbol():
    from rag_docvqa_b200 import build
    path = build.build()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    for name in declared_symbols():
        assert hasattr(lib, name), "librdv.so does not export %s" % name
    out = subprocess.run(["nm", "-D", "--defined-only", path], stdout=subprocess.PIPE, text=True).stdout
    exported = sorted(set(re.findall(r"\bT (rdv_\w+)", out)))
    assert exported == declared_symbols(), "exported %s != declared %s" % (exported, declared_symbols())


def test_binding_table_matches_header():
    from rag_docvqa_b200 import _lib
    assert sorted(_lib.SIGNATURES) == declared_symbols()
    assert _lib.lib.rdv_abi_version() == _lib.ABI_VERSION
    assert isinstance(_lib.lib.rdv_last_error(), bytes)      # "" in a fresh process; an earlier test may have provoked one


def test_sass_is_sm100a_only():
    from rag_docvqa_b200 import build
    out = subprocess.run(["cuobjdump", "--list-elf", build.LIB], stdout=subprocess.PIPE, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_argument_errors_are_reported_before_any_launch():
    """Every entry point validates its arguments first and returns a negative code with a message -- no CUDA call
    is made on that path, so it can be exercised without a GPU (error behaviour of the boundary, include/rdv.h:18)."""
    from rag_docvqa_b200 import _lib
    lib = _lib.lib
    cases = [
        (lib.rdv_rerank_order(None, 0, None, 4, 0, 0.4, 5, 1, None, None, None, None), _lib.E_LIMIT, "rerank_order"),      # k = 0
        (lib.rdv_rerank_order(None, 0, None, 4, 65, 0.4, 5, 1, None, None, None, None), _lib.E_LIMIT, "rerank_order"),     # k > 64
        (lib.rdv_rerank_order(None, 0, None, 4, 5, 0.4, -1, 1, None, None, None, None), _lib.E_INVALID, "negative"),
        (lib.rdv_rerank_order(None, 0, None, 4, 5, 0.4, 5, 1, None, None, None, None), _lib.E_INVALID, "null"),
        (lib.rdv_page_vote(None, None, None, None, 3, 0, 0, 1, None, None, None), _lib.E_LIMIT, "page_vote"),
        (lib.rdv_page_vote(None, None, None, None, 3, 5, 1, 1, None, None, None), _lib.E_INVALID, "null"),
        (lib.rdv_layout_assign(None, None, None, None, None, None, None, -1, 1, None, None, None, None), _lib.E_INVALID, "negative"),
        (lib.rdv_layout_assign(None, None, None, None, None, None, None, 3, 1, None, None, None, None), _lib.E_INVALID, "null"),
        (lib.rdv_topk_merge(None, None, 2, 4, 0, None, None, None), _lib.E_INVALID, "topk_merge"),
        (lib.rdv_s2_weights(None, None, -1, None, 0, 0, None, 4, None, None), _lib.E_INVALID, "negative"),
        (lib.rdv_s2_weights(None, None, 2, None, 0, 0, None, 4, None, None), _lib.E_INVALID, "null"),
        (lib.rdv_s2_weights(16, 16, 2, None, 0, 7, 16, 4, 16, None), _lib.E_INVALID, "what"),
        (lib.rdv_s2_weights(16, 16, 2, None, 0, _lib.S2_SEMANTIC, 16, 4, 16, None), _lib.E_INVALID, "embeddings"),
        (lib.rdv_s2_weights(8, 16, 2, None, 0, _lib.S2_SPATIAL, 16, 4, 16, None), _lib.E_ALIGN, "aligned"),
    ]
    for code, want, needle in cases:
        assert code == want, (code, want, needle)
    assert lib.rdv_layout_assign(None, None, None, None, None, None, None, 3, 1, None, None, None, None) == _lib.E_INVALID
    assert b"layout_assign" in lib.rdv_last_error()
    try:
        _lib.check(lib.rdv_rerank_order(None, 0, None, 4, 65, 0.4, 5, 1, None, None, None, None))
        raise AssertionError("no exception")
    except _lib.RdvError as exc:
        assert exc.code == _lib.E_LIMIT and "rerank_order" in str(exc)
    # empty batches are accepted and launch nothing
    assert lib.rdv_rerank_order(None, 0, None, 0, 5, 0.4, 5, 1, None, None, None, None) == _lib.OK
    assert lib.rdv_page_vote(None, None, None, None, 0, 5, 1, 1, None, None, None) == _lib.OK
    assert lib.rdv_layout_assign(None, None, None, None, None, None, None, 0, 1, None, None, None, None) == _lib.OK
    assert lib.rdv_s2_weights(None, None, 0, None, 0, 0, None, 0, None, None) == _lib.OK


def test_ctypes_struct_layouts_match_the_header():
    """Every struct the binding mirrors by hand has the size the compiler gave it (rdv_struct_size)."""
    from rag_docvqa_b200 import _lib
    mirrors = {"rdv_docstore": _lib.DocStoreStruct, "rdv_gather_args": _lib.GatherArgsStruct, "rdv_pagestore": _lib.PageStoreStruct,
               "rdv_visual_args": _lib.VisualArgsStruct, "rdv_p2s_args": _lib.P2SArgsStruct,
               "rdv_small_layout": _lib.SmallLayoutStruct, "rdv_vt5_embed_tables": _lib.EmbedTablesStruct}
    for name, mirror in mirrors.items():
        assert _lib.lib.rdv_struct_size(name.encode()) == ctypes.sizeof(mirror), name
    assert _lib.lib.rdv_struct_size(b"rdv_chunk_rec") == 32 and _lib.lib.rdv_struct_size(b"rdv_tok_rec") == 32
    assert _lib.lib.rdv_struct_size(b"no_such_struct") == -1


def test_header_is_plain_c_and_links_against_the_library(tmp_path):
    """include/rdv.h is what a cgo / JNI / N-API binding would include: it must compile as C99 (and as C++) on its own, and a
    C program that calls through it must link against librdv.so and agree with the binding about the ABI version."""
    import shutil
    from rag_docvqa_b200 import _lib, build
    gcc = shutil.which("gcc")
    if gcc is None:
        import pytest
        pytest.skip("no C compiler")
    lib = build.build()
    src = tmp_path / "abi.c"
    src.write_text('#include <stdio.h>\n#include "rdv.h"\n'
                   'int main(void) { printf("%d %lld\\n", rdv_abi_version(), (long long)rdv_struct_size("rdv_gather_args")); return 0; }\n')
    inc = os.path.join(ROOT, "include")
    subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-I", inc, str(src)], check=True)
    gxx = shutil.which("g++")
    if gxx:
        subprocess.run([gxx, "-std=c++17", "-Wall", "-Werror", "-fsyntax-only", "-x", "c++", "-I", inc, str(src)], check=True)
    exe = tmp_path / "abi"
    libdir = os.path.dirname(lib)
    subprocess.run([gcc, "-std=c99", "-I", inc, str(src), "-o", str(exe), "-L", libdir, "-l:librdv.so", "-Wl,-rpath," + libdir], check=True)
    out = subprocess.run([str(exe)], stdout=subprocess.PIPE, text=True, check=True).stdout.split()
    assert int(out[0]) == _lib.ABI_VERSION
    assert int(out[1]) == ctypes.sizeof(_lib.GatherArgsStruct)
