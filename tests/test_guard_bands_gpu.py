"""Out-of-bounds WRITE check of our own (compute-sanitizer is closed on this pool: profiles/r2a_sanitizer_unavailable.log).

Every CUDA tensor the product allocates while a test body runs -- outputs, workspaces, staging blobs -- is carved out of a
larger buffer with 4 KB sentinel bands on both sides; after the body (and a device synchronise) every band must be intact.
The bodies are the small parity cases of the other GPU test modules (which also compare every result with the oracle), at
deliberately awkward sizes: rows that do not fill a tile, d = 4, k > n, empty documents, documents above the selection
cache, ragged strips.  What this cannot see: out-of-bounds READS and races (the parity tests' bit-for-bit repeatability
checks are the evidence there)."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

BAND = 4096
FILL = 0xA5


class GuardedAllocations:
    def __init__(self):
        self.real = {n: getattr(torch, n) for n in ("empty", "zeros", "full", "empty_like", "zeros_like")}
        self.bufs = []

    @staticmethod
    def _is_cuda(device):
        if device is None:
            return False
        return torch.device(device).type == "cuda"

    def _carve(self, size, dtype, device):
        dtype = dtype or torch.get_default_dtype()
        if len(size) == 1 and isinstance(size[0], (tuple, list, torch.Size)):
            size = tuple(size[0])
        numel = 1
        for s in size:
            numel *= int(s)
        nbytes = numel * torch.empty((), dtype=dtype).element_size()
        pad = (-nbytes) % 256
        buf = self.real["full"]((BAND + nbytes + pad + BAND,), FILL, dtype=torch.uint8, device=device)
        self.bufs.append((buf, nbytes))
        return buf[BAND:BAND + nbytes].view(dtype).view(tuple(int(s) for s in size))

    def __enter__(self):
        me = self

        def empty(*size, dtype=None, device=None, pin_memory=False, **kw):
            if me._is_cuda(device) and not pin_memory and not kw:
                return me._carve(size, dtype, device)
            return me.real["empty"](*size, dtype=dtype, device=device, pin_memory=pin_memory, **kw)

        def zeros(*size, dtype=None, device=None, **kw):
            if me._is_cuda(device) and not kw:
                return me._carve(size, dtype, device).zero_()
            return me.real["zeros"](*size, dtype=dtype, device=device, **kw)

        def full(size, value, dtype=None, device=None, **kw):
            if me._is_cuda(device) and not kw:
                if dtype is None:
                    dtype = torch.float32 if isinstance(value, float) else torch.int64
                return me._carve((size,) if isinstance(size, int) else tuple(size), dtype, device).fill_(value)
            return me.real["full"](size, value, dtype=dtype, device=device, **kw)

        def empty_like(t, **kw):
            if t.is_cuda and not kw:
                return me._carve(tuple(t.shape), t.dtype, t.device)
            return me.real["empty_like"](t, **kw)

        def zeros_like(t, **kw):
            if t.is_cuda and not kw:
                return me._carve(tuple(t.shape), t.dtype, t.device).zero_()
            return me.real["zeros_like"](t, **kw)
        for name, fn in (("empty", empty), ("zeros", zeros), ("full", full), ("empty_like", empty_like), ("zeros_like", zeros_like)):
            setattr(torch, name, fn)
        return self

    def __exit__(self, *exc):
        for name, fn in self.real.items():
            setattr(torch, name, fn)

    def check(self):
        torch.cuda.synchronize()
        for buf, nbytes in self.bufs:
            head, tail = buf[:BAND], buf[BAND + nbytes:]
            assert bool((head == FILL).all()), "a kernel wrote BEFORE a %d-byte buffer" % nbytes
            assert bool((tail == FILL).all()), "a kernel wrote PAST a %d-byte buffer" % nbytes
        return len(self.bufs)


def run_guarded(fn, *args, **kwargs):
    with GuardedAllocations() as g:
        fn(*args, **kwargs)
    return g.check()


GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_guard_detects_an_overrun():
    with GuardedAllocations() as g:
        t = torch.empty((10,), dtype=torch.float32, device="cuda:0")
    base = g.bufs[0][0]
    assert t.data_ptr() == base.data_ptr() + BAND
    g.check()
    base[BAND + 40] = 0                              # one byte past the 40-byte buffer
    with pytest.raises(AssertionError, match="PAST"):
        g.check()


@pytest.mark.parametrize("algo", [1, 2, 3])
def test_score_topk_paths(algo):
    import test_score_topk_gpu as T
    for name in T.TEXT_CASES:
        assert run_guarded(T.test_golden_cases, GOLDEN, name, algo) > 0
    run_guarded(T.test_k_sweep_with_duplicates, 5, algo)
    run_guarded(T.test_k_sweep_with_duplicates, 20, algo)
    run_guarded(T.test_c3_slice_large_docs, algo)


def test_score_topk_shapes_and_special_values():
    import test_score_topk_gpu as T
    for d in (4, 100, 384, 640, 1024):
        run_guarded(T.test_dims, d)
    run_guarded(T.test_special_values)
    run_guarded(T.test_all_empty_and_zero_question)
    for B in (1, 3, 97):
        run_guarded(T.test_cluster_kernel_equals_two_launches, B)


def test_gather_and_packed_inputs():
    import test_retriever_gpu as T
    for s, reorder, sep in ((0, False, False), (0, True, True), (3, False, True), (200, True, False)):
        run_guarded(T.test_packed_inputs_vs_oracle_c2_slice, s, reorder, sep, True)
    run_guarded(T.test_packed_inputs_vs_oracle_c2_slice, 7, True, True, False)
    run_guarded(T.test_one_launch_step_equals_two_launches, True, True, 5, True)
    run_guarded(T.test_one_launch_step_equals_two_launches, True, False, 20, False)
    run_guarded(T.test_retrieve_c2_slice_vs_oracle, 0, False, "device")


def test_pool_maxsim_merge_pooled():
    import test_pool_maxsim_merge_gpu as T
    run_guarded(T.test_mean_pooling_golden, GOLDEN)
    run_guarded(T.test_mean_pooling_weighted_mask_and_normalise)
    for mode in ("ffma", "tf32x3"):
        run_guarded(T.test_late_interaction_golden, GOLDEN, mode)
    run_guarded(T.test_topk_segments_matches_oracle)
    run_guarded(T.test_topk_merge_equals_unsharded, 4, 10)
    run_guarded(T.test_pooled_patch_golden, GOLDEN)
    run_guarded(T.test_pooled_patch_c4_shape_and_nan)
    for split, n, L, d in (("auto", 8, 2048, 768), ("16", 3, 1000, 1024), ("2", 5, 70, 768), ("16", 1, 4096, 64)):
        with pytest.MonkeyPatch.context() as mp:                 # the cluster-split pooling kernel: leader's DSMEM rows
            run_guarded(T.test_mean_pooling_row_split_over_a_cluster, mp, split, n, L, d)


def test_input_embeddings():
    import test_vt5_embed_gpu as T
    run_guarded(T.test_golden_spatial_module_and_prepare_inputs, GOLDEN)
    for D, B, L in ((768, 5, 131), (1024, 2, 77), (100, 2, 33), (4, 1, 5), (768, 1, 1)):
        assert run_guarded(T.test_shapes_against_the_float64_oracle, D, B, L) > 0
    run_guarded(T.test_trimmed_views_of_the_gather_buffers_and_empty_batches)
    run_guarded(T.test_out_of_range_indices_are_flagged_not_read)
    run_guarded(T.test_after_the_gather_kernel)


def test_tensor_core_paths(monkeypatch):
    import test_tc_gpu as T
    for ctas in ("1", "2"):
        monkeypatch.setenv("RDV_TC_CTAS", ctas)
        run_guarded(T.test_corpus_topk_matches_exact_bf16_math, 300, 72, 5, 3)
        run_guarded(T.test_corpus_topk_matches_exact_bf16_math, 5000, 768, 300, 10)
        run_guarded(T.test_maxsim_bf16_tc, 2, 77, 513, 96)
    run_guarded(T.test_corpus_sharded_equals_unsharded)
    run_guarded(T.test_corpus_searcher_graph_equals_search_local)
    run_guarded(T.test_corpus_searcher_empty_shard)


def test_widening_rows():
    import test_chunker_gpu as TC
    import test_postproc_gpu as TP
    import test_s2chunker_gpu as TS
    import test_visual_pack_gpu as TV
    import test_pix2struct_gpu as TX
    ran = guarded = 0
    for mod in (TP, TC, TS, TV, TX):
        for name in sorted(dir(mod)):
            fn = getattr(mod, name)
            if not (name.startswith("test_") and callable(fn)):
                continue
            code = fn.__code__
            params = code.co_varnames[:code.co_argcount]
            if params == ("golden_dir",):
                guarded += run_guarded(fn, GOLDEN)
                ran += 1
            elif params == ():
                guarded += run_guarded(fn)
                ran += 1
    assert ran >= 8 and guarded > 50
