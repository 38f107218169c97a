#!/usr/bin/env python
"""bench.py -- retrieval throughput of the B200 path on the workload BASELINE.json names.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload C2|C3]
    (N > 1: launched by torchrun, one rank per GPU)

One "step" = the retrieval of ONE batch (B questions, one document each) = what Retriever.retrieve
does (reference src/_modules.py:2155-2180): cosine score of every chunk, per-document top-k, gather
of the hits.  Default workload: C2 = BASELINE.json configs[1] (64 questions x docs of <= 20 pages,
30 chunks/page, 384-d, k=5).

  value     device-resident: the two kernels of the step (streaming score; per-document top-k + gather into
            the generator's input_ids/boxes/mask) launched through the C ABI, inputs already in HBM.
            Successive steps rotate over R distinct resident batches (> 2x the 126 MB L2 in total), so
            every step streams its embeddings from HBM.  The K timed steps are captured once into a CUDA
            graph (2K kernel nodes) and the timed region is one launch of it; --lanes L (default 8) lets
            successive, independent batches rotate over L captured streams -- what a serving loop with L
            batches in flight does -- and --lanes 1 keeps one dependent chain.  Both, and the same steps as plain stream launches, are reported in `stages`.
  roofline  the dominant kernel (score_ldg_kernel) timed alone over the same rotation with CUDA
            events; achieved = algorithmic bytes per launch / mean launch duration.
  e2e       the drop-in `Retriever.retrieve` (reference signature: pinned HOST embeddings + the
            reference's nested lists + PIL pages in, the reference's 9-tuple out), H2D and D2H inside
            the timed region.  Patch crops are returned as deferred PIL images (rectangle computed, pixels
            cut on first use); the reference arm / cpu_baseline likewise computes the crop rectangles and
            skips the PIL pixel copy, so both arms time the same work.  Numbers WITH the eager PIL copy
            are reported for both under extras.
  reference arm (--impl reference): the oracle restatement of the reference's CPU path
            (oracle/ref_restated.py: same torch CPU ops, same Python list walk) on all host threads.

Prints ONE JSON line (rank 0).  DESIGN.md section "Measurement" derives every field.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
import zlib

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

L2_BYTES = 126 * 1024 * 1024
METRIC = "retrieval_queries_per_sec"


def ncu_traffic(key):
    """Per-launch DRAM traffic of a kernel from the committed ncu capture (profiles/traffic.json), or None."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        with open(path) as f:
            return json.load(f).get(key)
    except Exception:
        return None


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), float(p.get("bf16_tflops", 1590.0)), "measured"
    return 6650.0, 1590.0, "fallback"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""
    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu_index = gpu_index
        self.samples = []
        self._stop = threading.Event()
        self._thread = None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.gpu_index), "--query-gpu=" + self.QUERY,
                                      "--format=csv,noheader,nounits"], stdout=subprocess.PIPE,
                                     stderr=subprocess.DEVNULL, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self._stop.wait(0.05)

    def __enter__(self):
        self._thread = threading.Thread(target=self._run, daemon=True)
        self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        self._thread.join(timeout=6)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm = sorted(float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[3 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.samples[0][1]),
                "reasons": reasons, "samples": len(self.samples)}


def dist_env():
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def score_bytes(sizes, d, k):
    """N*d*4 (embeddings, read once) + B*d*4 (questions) + N*4 (every similarity
    written) + B*k*8 (top-k idx+val)."""
    n, b = int(sum(sizes)), len(sizes)
    return n * d * 4 + b * d * 4 + n * 4 + b * k * 8


def workload_text(w):
    return "%s: %d questions x docs of <=%d pages, %d chunks/page, %d-d, top-k=%d" % (
        w.name, w.docs, w.max_pages, w.chunks_per_page, w.dim, w.k)


def prompts_for(n):
    return [[5 + (zlib.crc32(t.encode()) % 1000) for t in ("question: what is item %d about ?  context: " % b).split()]
            for b in range(n)]


# ------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the oracle restatement of the reference's own CPU path
# ------------------------------------------------------------------------------------------------
def cpu_retrieve_fn(batch, k, crop):
    from oracle import ref_restated as R
    args = (batch["text_embeddings"], batch["question_embeddings"], batch["words_text_chunks"],
            batch["words_box_chunks"], batch["layout_labels_chunks"], batch["images"], batch["page_indices"])
    return lambda: R.retrieve(*args, k=k, crop=crop)


def cpu_score_topk_fn(batch, k):
    from oracle import ref_restated as R
    emb, q = batch["text_embeddings"], batch["question_embeddings"]

    def fn():
        sims = R.score(emb, q)
        return [R.topk_reference(s, k) for s in sims]
    return fn


def time_cpu(fn, seconds, min_reps=2):
    fn()
    best, reps, t_end = float("inf"), 0, time.perf_counter() + seconds
    while reps < min_reps or time.perf_counter() < t_end:
        t0 = time.perf_counter()
        fn()
        best = min(best, time.perf_counter() - t0)
        reps += 1
    return best, reps


def run_reference(args):
    from rag_docvqa_b200 import synth
    rank, world, _ = dist_env()
    if rank != 0:
        return
    if args.workload == "C5":
        return run_reference_corpus(args)
    if args.workload == "C4":
        return run_reference_visual(args)
    w = synth.WORKLOADS[args.workload]
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    batch = synth.make_text_batch(args.workload, with_lists=True, share_image_pool=24)
    step = cpu_retrieve_fn(batch, w.k, crop=False)
    for _ in range(max(1, min(args.warmup, 3))):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    qps = w.docs * args.steps / dt
    sample = ("oracle/ref_restated.py retrieve() = score + torch.topk + Python list gather + compact + crop "
              "rectangles (PIL pixel copy excluded, as in the GPU arm) on one full %s batch per step, %d steps"
              % (w.name, args.steps))
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": qps, "unit": "queries/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_text(w)},
        "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


def run_reference_corpus(args):
    """CPU arm of the corpus mode: torch matmul + topk (what the reference's scoring ops give for Q
    questions against N chunks) on a 1/64 row slice per step, time scaled x64."""
    from oracle import ref_restated as R
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    N, d, Qn, k = args.corpus_rows, 768, args.corpus_queries, 10
    n_cpu = max(1, N // 64)
    g = torch.Generator().manual_seed(synth_seed(5))
    e = (torch.randn(n_cpu, d, generator=g) / d ** 0.5).to(torch.bfloat16).float()
    q = torch.randn(Qn, d, generator=g) / d ** 0.5

    def step():
        return torch.topk(R.corpus_scores(e, q), k, dim=1)
    for _ in range(max(1, min(args.warmup, 2))):
        step()
    steps = max(1, min(args.steps, 10))
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps * 64
    qps = Qn / dt
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": qps, "unit": "queries/s", "n_gpus": args.gpus, "steps": steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "C5: %d chunks x %d-d, %d questions, top-k=%d" % (N, d, Qn, k)},
        "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": threads, "kind": "port",
                         "sample": "oracle corpus_scores + torch.topk on a 1/64 row slice (%d rows) per step, time x64" % n_cpu},
        "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


def run_reference_visual(args):
    """CPU arm of the visual path: the oracle's late_interaction (the reference's torch ops) on 5 of the 50 strips
    of one question per step, time scaled x10."""
    from oracle import ref_restated as R
    from rag_docvqa_b200 import synth
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    strips, L, d, n_cpu = 50, 2048, 768, 5
    patches, q = synth.make_strip_batch(1, [n_cpu], L, d, synth_seed(4))
    for _ in range(max(1, min(args.warmup, 2))):
        R.late_interaction(q[0:1], patches[0])
    steps = max(1, min(args.steps, 10))
    t0 = time.perf_counter()
    for _ in range(steps):
        R.late_interaction(q[0:1], patches[0])
    dt = (time.perf_counter() - t0) / steps * (strips / n_cpu)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": 1.0 / dt, "unit": "queries/s", "n_gpus": args.gpus, "steps": steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "C4: questions x %d strips x (%d x %d) tokens, MaxSim late interaction" % (strips, L, d)},
        "cpu_baseline": {"value": 1.0 / dt, "unit": "queries/s", "cores": threads, "kind": "port",
                         "sample": "oracle late_interaction on %d of %d strips of one question per step, time x%d" % (n_cpu, strips, strips // n_cpu)},
        "e2e": {"value": 1.0 / dt, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def timed_loop(fn, steps, barrier):
    """`steps` calls of fn(i) between two CUDA events on the current stream; returns ms total."""
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for i in range(steps):
        fn(i)
    ev1.record()
    barrier()
    return ev0.elapsed_time(ev1)


def stage_extras(dev, hbm_peak, steps):
    """Per-stage numbers for the other kernels of the path (rank 0, N=1 only)."""
    from rag_docvqa_b200 import functional as F
    from rag_docvqa_b200 import synth
    out = {}
    sync = torch.cuda.synchronize

    # masked mean pooling on a C2-shaped token batch: 8192 chunks x ~96 of <=160 tokens x 384
    embs, mask = synth.make_token_batch(8192, 384, 7, device=dev, max_len=160)
    valid = int(mask.sum().item())
    n, L, d = embs.shape
    bytes_pool = valid * d * 4 + n * L * 8 + n * d * 4
    for _ in range(3):
        F.mean_pooling(embs, mask)
    ms = timed_loop(lambda i: F.mean_pooling(embs, mask), 20, sync) / 20
    out["mean_pool_f32"] = {"shape": [n, L, d], "valid_tokens": valid, "ms": ms, "algorithmic_bytes": bytes_pool,
                            "GBps": bytes_pool / ms / 1e6, "frac_hbm": bytes_pool / ms / 1e6 / hbm_peak,
                            "note": "tensor (%.0f MB) > L2" % (embs.numel() * 4 / 1e6)}
    del embs, mask

    # MaxSim fp32 (FFMA) on one C4-shaped document: 50 strips x 2048 x 768 vs a 2048 x 768 question
    patches, q = synth.make_strip_batch(1, [50], 2048, 768, 3, device=dev)
    flops = 2.0 * 50 * 2048 * 2048 * 768
    for _ in range(2):
        F.late_interaction(q[0:1], patches[0], mode="ffma")
    ms = timed_loop(lambda i: F.late_interaction(q[0:1], patches[0], mode="ffma"), 5, sync) / 5
    out["maxsim_f32"] = {"shape": "50 strips x 2048 x 768 vs 2048 x 768", "ms": ms, "TFLOPs": flops / ms / 1e9,
                         "bound": "CUDA-core FFMA (fp32 parity mode)", "questions_per_s": 1e3 / ms}
    # the same contraction at fp32 grade on the tensor pipe: 3 tf32 products per fp32 product (split included)
    for _ in range(2):
        F.late_interaction(q[0:1], patches[0], mode="tf32x3")
    ms = timed_loop(lambda i: F.late_interaction(q[0:1], patches[0], mode="tf32x3"), 10, sync) / 10
    out["maxsim_tf32x3_tc"] = {"shape": "50 strips x 2048 x 768 vs 2048 x 768", "ms": ms, "fp32_equivalent_TFLOPs": flops / ms / 1e9,
                               "tf32_TFLOPs_issued": 3 * flops / ms / 1e9, "bound": "tensor (tcgen05 kind::tf32, 3 products per fp32 product)",
                               "questions_per_s": 1e3 / ms, "includes": "normalise + hi/lo split of Q and P"}
    # the same document through the bf16 tcgen05 mode (normalise + cast included)
    _, tf_peak, _ = measured_peaks()
    for _ in range(2):
        F.late_interaction_bf16(q[0:1], patches[0])
    ms = timed_loop(lambda i: F.late_interaction_bf16(q[0:1], patches[0]), 10, sync) / 10
    out["maxsim_bf16_tc"] = {"shape": "50 strips x 2048 x 768 vs 2048 x 768", "ms": ms, "TFLOPs": flops / ms / 1e9,
                             "frac_tensor": flops / ms / 1e9 / tf_peak, "bound": "tensor (tcgen05, fp32 accumulate in TMEM)",
                             "questions_per_s": 1e3 / ms, "includes": "fp32->normalised bf16 cast of Q and P"}
    del patches, q

    # corpus mode at C5's per-rank shape: 1.25 M x 768 bf16 rows, 1024 questions, k=10
    from rag_docvqa_b200 import sharded
    g = torch.Generator(device=dev)
    g.manual_seed(5)
    rows = torch.empty((1_250_000, 768), dtype=torch.bfloat16, device=dev)
    for a in range(0, rows.shape[0], 250_000):
        rows[a:a + 250_000] = torch.randn(250_000, 768, generator=g, device=dev).to(torch.bfloat16)
    shard = sharded.CorpusShard(rows)
    qs = torch.randn(1024, 768, generator=g, device=dev)
    for _ in range(2):
        shard.search_local(qs, 10)
    ms = timed_loop(lambda i: shard.search_local(qs, 10), 10, sync) / 10
    cflops = 2.0 * 1024 * rows.shape[0] * 768
    out["corpus_bf16_tc_per_rank_C5"] = {"shape": "1024 questions x 1.25 M x 768 bf16, k=10", "ms": ms,
                                         "TFLOPs": cflops / ms / 1e9, "frac_tensor": cflops / ms / 1e9 / tf_peak,
                                         "questions_per_s": 1024 / ms * 1e3}
    del rows, shard, qs

    # score+top-k on C3 (long documents, 768-d, k=10): far above L2, shows the kernel's streaming rate
    w3 = synth.WORKLOADS["C3"]
    b3 = synth.make_text_batch("C3", device=dev)
    t3 = F.build_doc_table(b3["text_embeddings"], w3.dim, dev)
    for _ in range(3):
        F.score_topk_table(t3, b3["question_embeddings"], w3.k)
    ms = timed_loop(lambda i: F.score_topk_table(t3, b3["question_embeddings"], w3.k), steps, sync) / steps
    by = score_bytes(b3["sizes"], w3.dim, w3.k)
    out["score_topk_f32_C3"] = {"workload": workload_text(w3), "ms": ms, "algorithmic_bytes": by,
                                "GBps": by / ms / 1e6, "frac_hbm": by / ms / 1e6 / hbm_peak,
                                "queries_per_s": w3.docs / ms * 1e3}
    return out


def run_ours(args):
    from rag_docvqa_b200 import functional as F
    from rag_docvqa_b200 import synth, _lib
    from rag_docvqa_b200.docstore import DocStore
    from rag_docvqa_b200.retriever import Retriever
    rank, world, local = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    w = synth.WORKLOADS[args.workload]
    hbm_peak, _, peak_kind = measured_peaks()
    warmup = max(3, args.warmup)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world > 1:
            t = torch.tensor([x], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return x

    # ---- setup (untimed): lists + DocStore once, R resident embedding batches -------------------------
    with_lists = args.workload != "C3"
    # weak scaling: every rank gets documents of the SAME sizes (the seeded C2 / C3 shape) with its own embedding
    # values, so per-GPU work is fixed as N grows
    base_seed = synth.SEED_BASE + w.config_id
    rank_seed = base_seed + 100000 * rank
    host_batch = synth.make_text_batch(args.workload, with_lists=with_lists, share_image_pool=24, seed=base_seed)
    sizes = host_batch["sizes"]
    step_bytes = score_bytes(sizes, w.dim, w.k)
    R = max(2, min(16, int(np.ceil(2.2 * L2_BYTES / max(1, step_bytes)))))
    while R % max(1, args.lanes):     # a batch's buffers are only ever touched by one lane
        R += 1
    batches = [synth.make_text_batch(args.workload, device=dev, seed=base_seed, emb_seed=rank_seed + 1000 * (r + 1))
               for r in range(R)]
    tables = [F.build_doc_table(b["text_embeddings"], w.dim, dev, algo=args.algo) for b in batches]
    outs = [dict(sims=torch.empty(t.total_rows, dtype=torch.float32, device=dev),
                 idx=torch.empty((t.B, w.k), dtype=torch.int32, device=dev),
                 val=torch.empty((t.B, w.k), dtype=torch.float32, device=dev),
                 cnt=torch.empty((t.B,), dtype=torch.int32, device=dev)) for t in tables]
    stream = torch.cuda.current_stream(dev).cuda_stream
    # a step = streaming score kernel (rdv_score_f32) + ONE kernel per batch that selects each document's
    # top-k and gathers it into the generator tensors (rdv_gather_vt5_inputs with fused selection).  Without a
    # DocStore (C3) the second kernel is the stand-alone selection (rdv_topk_segments_f32).
    lib = _lib.lib
    stream_algo = [(_lib.SCORE_LDG if t.algo == _lib.SCORE_LDG_FUSED else t.algo) for t in tables]
    score_args, select_args = [], []
    for t, o, b, algo in zip(tables, outs, batches, stream_algo):
        p_tiles, p_row = t.pointers()
        score_args.append((p_tiles, t.total_tiles, t.tile_rows, algo, b["question_embeddings"].data_ptr(), t.B, t.d,
                           o["sims"].data_ptr(), stream))
        select_args.append((o["sims"].data_ptr(), p_row, t.B, w.k, t.max_rows, o["idx"].data_ptr(), o["val"].data_ptr(),
                            o["cnt"].data_ptr(), stream))
    plans = None
    if with_lists:
        table = synth.make_tokens_for_words(host_batch["words_text_chunks"], seed=3)
        t_store = time.perf_counter()
        store = DocStore.from_lists(host_batch["words_text_chunks"], host_batch["words_box_chunks"],
                                    host_batch["layout_labels_chunks"], host_batch["page_indices"],
                                    lambda wd: table.get(wd, [2]), dev, images=host_batch["images"])
        docstore_build_s = time.perf_counter() - t_store
        prompts = prompts_for(w.docs)
        plans = [store.prepare_gather(o["idx"], o["cnt"], prompts, max_len=512, sims=o["sims"], topk_val=o["val"],
                                      max_rows=t.max_rows) for o, t in zip(outs, tables)]
    torch.cuda.synchronize()

    def make_launchers(stream_ptr):
        """(score, gather, step) closures launching on the given stream through the C ABI."""
        s_args = [a[:-1] + (stream_ptr,) for a in score_args]
        g_args = [a[:-1] + (stream_ptr,) for a in select_args]

        def score(i):
            rc = lib.rdv_score_f32(*s_args[i % R])
            if rc:
                _lib.check(rc)

        def gather(i):
            if plans is not None:
                plans[i % R].launch(stream_ptr)
            else:
                rc = lib.rdv_topk_segments_f32(*g_args[i % R])
                if rc:
                    _lib.check(rc)

        def step(i):
            score(i)
            gather(i)
        return score, gather, step

    launch_score, launch_gather, launch_step = make_launchers(stream)
    launches_per_step = 2
    for i in range(warmup):
        launch_step(i)
    torch.cuda.synchronize()

    # The timed region is ONE CUDA-graph launch holding exactly `steps` steps (2 kernel nodes each), rotating over
    # the R resident batches: the hot loop is launch-bound (a step is ~13 us of device time, ~8 us of host
    # enqueue), so it is captured once and replayed.  `lanes` = 1: the steps form one dependent chain.
    # `lanes` = L: successive (independent) batches rotate over L captured streams, so the latency-bound
    # select+gather of one batch overlaps the HBM-bound score of the next, as a serving loop with L batches in
    # flight would (measured on B200, C2: 13.0 us per step with 1 lane, 11.7 with 2, 8.1 with 3, 7.2 with 4, 5.1 with 8 --
    # 32.5 MB per 5.1 us = 6.3 TB/s: the whole step at the HBM roofline).
    def capture(fn, n, lanes=1):
        g = torch.cuda.CUDAGraph()
        main = torch.cuda.Stream(dev)
        extra = [torch.cuda.Stream(dev) for _ in range(lanes - 1)]
        fns = [fn(st.cuda_stream) for st in [main] + extra]
        main.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.graph(g, stream=main):
            for st in extra:
                st.wait_stream(main)
            for i in range(n):
                fns[i % lanes](i)
            for st in extra:
                main.wait_stream(st)
        return g

    def timed_graph(g, reps):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record()
        for _ in range(reps):
            g.replay()
        ev1.record()
        barrier()
        return ev0.elapsed_time(ev1)

    g_seq = capture(lambda sp: make_launchers(sp)[2], args.steps, 1)
    g_pipe = capture(lambda sp: make_launchers(sp)[2], args.steps, max(2, args.lanes))
    g_score = capture(lambda sp: make_launchers(sp)[0], args.steps, 1)
    g_gather = capture(lambda sp: make_launchers(sp)[1], args.steps, 1)
    for g in (g_seq, g_pipe, g_score, g_gather):
        g.replay()
    torch.cuda.synchronize()

    # ---- value: K steps, device resident --------------------------------------------------------------
    with ClockSampler(local) as clocks:
        ms_seq = timed_graph(g_seq, 1)
        ms_pipe = timed_graph(g_pipe, 1)
        # per-kernel timings over the same rotation (roofline = the dominant kernel alone)
        ms_score = timed_graph(g_score, 1) / args.steps
        ms_gather = timed_graph(g_gather, 1) / args.steps
        ms_plain = timed_loop(launch_step, args.steps, barrier)          # the same steps as plain stream launches
        ms_score_plain = timed_loop(launch_score, args.steps, barrier) / args.steps
        t_end = time.perf_counter() + 0.6          # keep the GPU busy so the sampler sees clocks under load
        while time.perf_counter() < t_end:
            g_pipe.replay()
            torch.cuda.synchronize()
    ms_seq, ms_pipe, ms_plain = max_over_ranks(ms_seq), max_over_ranks(ms_pipe), max_over_ranks(ms_plain)
    pipelined = args.lanes >= 2
    ms_total = ms_pipe if pipelined else ms_seq
    ms_per_step = ms_total / args.steps
    qps = w.docs * world / (ms_per_step * 1e-3)
    achieved = step_bytes / (ms_score * 1e-3) / 1e9

    # ---- e2e: the drop-in Retriever.retrieve with HOST inputs ------------------------------------------
    e2e = None
    extras = {}
    host_sets = [([e.cpu().pin_memory() for e in b["text_embeddings"]], b["question_embeddings"].cpu().pin_memory())
                 for b in batches[:min(R, 4)]]
    h2d = sum(e.numel() * 4 for e in host_sets[0][0]) + host_sets[0][1].numel() * 4
    cfg = {"compute_stats": False, "compute_stats_examples": False, "n_stats_examples": 0, "chunk_num": w.k,
           "device": str(dev)}
    e2e_steps = max(5, min(args.steps, 40))
    if args.skip_e2e:
        e2e = {"value": None, "unit": "queries/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 0,
               "api": "skipped (--skip-e2e: profiling run, keeps the kernel launch list to the device-resident steps)"}
    elif with_lists:
        lists = (host_batch["words_text_chunks"], host_batch["words_box_chunks"], host_batch["layout_labels_chunks"],
                 host_batch["images"], host_batch["page_indices"])
        retr = Retriever({**cfg, "retrieval_lazy_patches": True})

        def e2e_step(i):
            emb_h, q_h = host_sets[i % len(host_sets)]
            return retr.retrieve(emb_h, q_h, *lists)
        for i in range(3):
            e2e_step(i)
        barrier()
        t0 = time.perf_counter()
        for i in range(e2e_steps):
            out = e2e_step(i)
        torch.cuda.synchronize()
        e2e_dt = max_over_ranks(time.perf_counter() - t0)
        d2h = w.docs * (w.k + 1) * 4 + sum(sizes) * 4
        e2e = {"value": w.docs * world * e2e_steps / e2e_dt, "unit": "queries/s", "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "ms_per_step": e2e_dt / e2e_steps * 1e3,
               "api": "rag_docvqa_b200.retriever.Retriever.retrieve (reference signature; pinned host embeddings, "
                      "nested lists and PIL pages in; 9-tuple out; patches = deferred crops)"}
    else:
        def e2e_step(i):
            emb_h, q_h = host_sets[i % len(host_sets)]
            table_h = F.upload_doc_table(emb_h, w.dim, dev)
            res = F.score_topk_table(table_h, q_h.to(dev, non_blocking=True), w.k)
            return res.topk_idx.cpu(), res.topk_cnt.cpu()
        for i in range(2):
            e2e_step(i)
        barrier()
        e2e_steps = 5
        t0 = time.perf_counter()
        for i in range(e2e_steps):
            e2e_step(i)
        torch.cuda.synchronize()
        e2e_dt = max_over_ranks(time.perf_counter() - t0)
        e2e = {"value": w.docs * world * e2e_steps / e2e_dt, "unit": "queries/s", "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": w.docs * (w.k + 1) * 4, "ms_per_step": e2e_dt / e2e_steps * 1e3,
               "api": "rag_docvqa_b200.functional.upload_doc_table + score_topk_table (pinned host embeddings in, top-k out)"}

    line = {
        "metric": METRIC, "value": qps, "unit": "queries/s", "n_gpus": world, "steps": args.steps,
        "warmup": warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_text(w),
                   "step": "streaming score kernel + %s, one batch of %d questions; the %d steps are one CUDA-graph "
                           "launch, %s" % (
                       "select+gather kernel (per-document top-k, packed VT5 inputs, max_source_length 512)"
                       if plans is not None else "per-document top-k kernel", w.docs, args.steps,
                       "successive independent batches rotate over %d captured streams (the latency-bound select+gather "
                       "of one batch overlaps the HBM-bound score of the next ones)" % args.lanes if pipelined
                       else "one dependent chain"),
                   "l2": "inputs larger than L2: %d distinct resident batches rotated (%.0f MB in total)" % (
                       R, R * step_bytes / 1e6),
                   "parallelism": "documents sharded across ranks (dp%d), no data-path collective" % world},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                     "frac": achieved / hbm_peak, "peak_kind": peak_kind,
                     "traffic": ncu_traffic("%s:%s" % ("score_tma_kernel" if stream_algo[0] == 2 else "score_ldg_kernel", w.name))
                     if rank == 0 else None,
                     "kernel": "score_tma_kernel" if stream_algo[0] == 2 else "score_ldg_kernel", "algorithmic_bytes_per_launch": step_bytes,
                     "ms_per_launch": ms_score},
        "e2e": e2e,
        "gpu_launches": args.steps * launches_per_step,
        "clocks": clocks.summary(),
        "stages": {"score_ms": ms_score, "select_gather_ms": ms_gather, "step_ms": ms_per_step,
                   "step_ms_graph_one_chain": ms_seq / args.steps, "step_ms_graph_lanes": ms_pipe / args.steps,
                   "step_ms_plain_stream_launches": ms_plain / args.steps, "score_ms_plain_stream_launches": ms_score_plain,
                   "step_GBps": step_bytes / (ms_per_step * 1e-3) / 1e9, "step_frac_hbm": step_bytes / (ms_per_step * 1e-3) / 1e9 / hbm_peak},
    }

    if rank == 0 and world == 1:
        threads = os.cpu_count() or 1
        torch.set_num_threads(threads)
        cpu_batch = dict(host_batch)
        cpu_batch["text_embeddings"] = [e.cpu() for e in batches[0]["text_embeddings"]]
        cpu_batch["question_embeddings"] = batches[0]["question_embeddings"].cpu()
        st_best, st_reps = time_cpu(cpu_score_topk_fn(cpu_batch, w.k), min(3.0, args.cpu_seconds))
        if with_lists:
            best, reps = time_cpu(cpu_retrieve_fn(cpu_batch, w.k, crop=False), args.cpu_seconds)
            line["cpu_baseline"] = {
                "value": w.docs / best, "unit": "queries/s", "cores": threads, "kind": "port",
                "sample": "oracle retrieve() (score + torch.topk + Python list gather + crop rectangles, no PIL pixel "
                          "copy) on one full %s batch, best of %d reps" % (w.name, reps),
                "score_topk_only_queries_per_s": w.docs / st_best}
            if not args.no_extras:
                # the same comparison WITH the eager PIL pixel copy of the reference, both arms
                eager = Retriever(cfg)
                t0 = time.perf_counter()
                for i in range(2):
                    eager.retrieve(host_sets[i % len(host_sets)][0], host_sets[i % len(host_sets)][1], *lists)
                torch.cuda.synchronize()
                extras["e2e_eager_pil_crops_queries_per_s"] = w.docs * 2 / (time.perf_counter() - t0)
                best_c, _ = time_cpu(cpu_retrieve_fn(cpu_batch, w.k, crop=True), 2.0, min_reps=1)
                extras["cpu_eager_pil_crops_queries_per_s"] = w.docs / best_c
                # the reference's own situation: the embedder left the embeddings on the GPU (src/RAGVT5.py:230-252)
                dev_emb, dev_q = batches[0]["text_embeddings"], batches[0]["question_embeddings"]
                retr.retrieve(dev_emb, dev_q, *lists)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                for i in range(e2e_steps):
                    retr.retrieve(dev_emb, dev_q, *lists)
                torch.cuda.synchronize()
                extras["retrieve_device_embeddings_queries_per_s"] = w.docs * e2e_steps / (time.perf_counter() - t0)
                # B200-native API: host embeddings in, packed generator tensors on the device out
                def packed_step(i):
                    emb_h, q_h = host_sets[i % len(host_sets)]
                    packed, _ = retr.retrieve_packed(emb_h, q_h, store, prompts)
                    return packed
                packed_step(0)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                for i in range(e2e_steps):
                    packed_step(i)
                torch.cuda.synchronize()
                extras["e2e_packed_queries_per_s"] = w.docs * e2e_steps / (time.perf_counter() - t0)
                # the packed path needs the documents pre-tokenised (DocStore): a per-DOCUMENT cost paid at ingest, not per
                # question -- it walks every word of every chunk on the host
                extras["docstore_build_s_per_batch_of_documents"] = docstore_build_s
                extras["docstore_words"] = store.n_words
                # ... and with the visual input as well: crops of the hits grid-packed and resized to 224 x 224 on the device
                from rag_docvqa_b200.pagestore import PageStore
                pstore = PageStore.from_images(host_batch["images"], dev)

                def packed_visual_step(i):
                    emb_h, q_h = host_sets[i % len(host_sets)]
                    return retr.retrieve_packed(emb_h, q_h, store, prompts, pages=pstore)
                packed_visual_step(0)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                for i in range(e2e_steps):
                    packed_visual_step(i)
                torch.cuda.synchronize()
                extras["e2e_packed_with_visual_input_queries_per_s"] = w.docs * e2e_steps / (time.perf_counter() - t0)
                # the visual pack kernels alone (hits resident): crop + grid pack + Pillow-exact bicubic resize + normalise
                pk, rs = retr.retrieve_packed(batches[0]["text_embeddings"], batches[0]["question_embeddings"], store, prompts)
                vplan = pstore.prepare_pack(pk.hit_page, pk.hit_rect, rs.topk_cnt)
                for _ in range(3):
                    vplan.launch()
                ms_v = timed_loop(lambda i: vplan.launch(), 20, torch.cuda.synchronize) / 20
                area = int(((pk.hit_rect[..., 2] - pk.hit_rect[..., 0]).clamp(min=0) * (pk.hit_rect[..., 3] - pk.hit_rect[..., 1]).clamp(min=0)).sum().item())
                extras["visual_pack"] = {"ms": ms_v, "patch_pixels": area, "algorithmic_bytes": area * 3 + w.docs * 224 * 224 * 15,
                                         "GBps": (area * 3 + w.docs * 224 * 224 * 15) / ms_v / 1e6,
                                         "cpu_reference_ms": None,
                                         "what": "64 documents x 5 crops -> grid canvas -> 224 x 224 bicubic (Pillow-exact) -> fp32 pixel_values"}
                # the same on the host, as the reference does it (PIL crop + concatenate grid + PIL resize), one thread
                from oracle import ref_restated as R_
                hr, hp, hc = pk.hit_rect.cpu().numpy(), pk.hit_page.cpu().numpy(), rs.topk_cnt.cpu().numpy()
                from PIL import Image as _Image
                t0 = time.perf_counter()
                for b in range(min(w.docs, 16)):
                    patches = [host_batch["images"][b][int(hp[b, j])].crop(tuple(int(v) for v in hr[b, j])) for j in range(int(hc[b]))]
                    if not patches:
                        continue
                    gw, gh, pos = R_.grid_layout([p_.size for p_ in patches])
                    canvas = _Image.new("RGB", (gw, gh))
                    for p_, xy in zip(patches, pos):
                        canvas.paste(p_, xy)
                    canvas.resize((224, 224), resample=_Image.Resampling.BICUBIC)
                extras["visual_pack"]["cpu_reference_ms"] = (time.perf_counter() - t0) * 1e3 * w.docs / min(w.docs, 16)
                # what consumes the hits: reranker index list + re-emission of the packed inputs in that
                # order, and the page vote, all on the device (the cross-encoder is a model: random scores stand in for it)
                from rag_docvqa_b200 import postproc as _pp
                pk, rs, gplan = retr.retrieve_packed(batches[0]["text_embeddings"], batches[0]["question_embeddings"], store,
                                                     prompts, return_plan=True)
                ce_scores = torch.rand((w.docs, w.k), device=dev)
                row_off_d = torch.from_numpy(np.concatenate([[0], np.cumsum(rs.sizes)]).astype(np.int64)).to(dev)

                def rerank_step(i):
                    order, kept, _ = _pp.rerank_order(ce_scores, rs.topk_cnt, 0.4, 5, 1)
                    gplan.set_emit_order(order, kept)
                    gplan.launch()
                for _ in range(3):
                    rerank_step(0)
                extras["rerank_packed_ms"] = timed_loop(rerank_step, 20, torch.cuda.synchronize) / 20
                for _ in range(3):
                    _pp.page_vote(pk.hit_page, rs.topk_cnt, rs.sims, row_off_d, True)
                extras["page_vote_weighted_ms"] = timed_loop(
                    lambda i: _pp.page_vote(pk.hit_page, rs.topk_cnt, rs.sims, row_off_d, True), 20, torch.cuda.synchronize) / 20
                # what feeds the path: Chunker.get_chunks with layout boxes, word x box containment on
                # the device, against the oracle's Python loops (16 documents)
                from rag_docvqa_b200.chunker import Chunker as _Chunker
                cw, cb, ci = synth.make_chunker_batch(77, docs=16, max_pages=20, max_words=700, max_layouts=30, degenerate=False)
                chunker = _Chunker({**cfg, "page_retrieval": "concat"})
                chunker.get_chunks(cw[:2], cb[:2], ci[:2], question_id=[0, 1])
                t0 = time.perf_counter()
                got_chunks = chunker.get_chunks(cw, cb, ci, question_id=list(range(16)))
                t_gpu = time.perf_counter() - t0
                t0 = time.perf_counter()
                want_chunks, _ = R_.get_chunks(cw, cb, ci)
                t_cpu = time.perf_counter() - t0
                extras["chunker_get_chunks"] = {
                    "documents": 16, "pages": sum(len(d) for d in cw), "words": sum(len(p) for d in cw for p in d),
                    "word_x_layout_box_pairs": sum(len(p) * len(g["boxes"]) for d, gi in zip(cw, ci) for p, g in zip(d, gi)),
                    "s": t_gpu, "cpu_oracle_s": t_cpu, "identical": got_chunks[0] == want_chunks[0] and got_chunks[2] == want_chunks[2]}
        else:
            line["cpu_baseline"] = {"value": w.docs / st_best, "unit": "queries/s", "cores": threads, "kind": "port",
                                    "sample": "oracle score+topk on one full %s batch, best of %d reps" % (w.name, st_reps)}
        if not args.no_extras:
            del batches, tables, outs
            torch.cuda.empty_cache()
            extras.update(stage_extras(dev, hbm_peak, 20))
        line["extras"] = extras
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------
# corpus mode (BASELINE.json configs[4]): row-sharded bf16 corpus, tcgen05 scoring, NCCL all-gather + merge
# ------------------------------------------------------------------------------------------------
def run_corpus(args):
    from rag_docvqa_b200 import functional as F
    from rag_docvqa_b200 import sharded
    rank, world, local = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    _, tf_peak, peak_kind = measured_peaks()
    N, d, Qn, k = args.corpus_rows, 768, args.corpus_queries, 10
    lo, hi = sharded.shard_bounds(N, world, rank)
    g = torch.Generator(device=dev)
    g.manual_seed(synth_seed(5) + rank)
    rows = torch.empty((hi - lo, d), dtype=torch.bfloat16, device=dev)
    u = torch.randn(d, generator=torch.Generator(device=dev).manual_seed(77), device=dev)
    u = u / u.norm()
    for a in range(0, hi - lo, 1 << 20):                      # generated shard by shard on the device
        b = min(hi - lo, a + (1 << 20))
        rows[a:b] = (torch.randn(b - a, d, generator=g, device=dev) / d ** 0.5 + 0.5 * u).to(torch.bfloat16)
    shard = sharded.CorpusShard(rows, id_offset=lo)
    gq = torch.Generator(device="cpu").manual_seed(synth_seed(5))
    q_host = (torch.randn(Qn, d, generator=gq) / d ** 0.5 + 0.5 * u.cpu()).pin_memory()
    q_dev = q_host.to(dev)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world > 1:
            t = torch.tensor([x], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return x

    def step(i):
        return sharded.search(shard, q_dev, k)
    warmup = max(3, args.warmup)
    for i in range(warmup):
        step(i)
    with ClockSampler(local) as clocks:
        ms_total = timed_loop(step, args.steps, barrier)
        ms_kernel = timed_loop(lambda i: shard.candidates(q_dev, k), args.steps, barrier) / args.steps
        t_end = time.perf_counter() + 0.5
        while time.perf_counter() < t_end:
            step(0)
            torch.cuda.synchronize()
    ms_per_step = max_over_ranks(ms_total) / args.steps
    ms_kernel = max_over_ranks(ms_kernel)
    flops = 2.0 * Qn * (hi - lo) * d

    def e2e_step(i):
        val, idx = sharded.search(shard, q_host.to(dev, non_blocking=True), k)
        return val.cpu(), idx.cpu()
    for i in range(2):
        e2e_step(i)
    barrier()
    e2e_steps = max(5, min(args.steps, 20))
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        out = e2e_step(i)
    torch.cuda.synchronize()
    e2e_dt = max_over_ranks(time.perf_counter() - t0)
    line = {
        "metric": METRIC, "value": Qn / (ms_per_step * 1e-3), "unit": "queries/s", "n_gpus": world, "steps": args.steps,
        "warmup": warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "C5: %d chunks x %d-d bf16 row-sharded over %d GPU(s), %d questions, top-k=%d" % (N, d, world, Qn, k),
                   "step": "bf16 cast of the questions, tcgen05 score + fused top-k on the local shard, local merge, "
                           "NCCL all-gather of (Q,k) candidates, final merge",
                   "l2": "inputs larger than L2 (%.1f GB per rank)" % ((hi - lo) * d * 2 / 1e9),
                   "parallelism": "corpus rows sharded across ranks (dp%d), one all-gather of %d bytes per rank" % (
                       world, Qn * k * 12)},
        "roofline": {"bound": "tensor", "achieved": flops / (ms_kernel * 1e-3) / 1e12, "peak": tf_peak, "unit": "TFLOP/s",
                     "frac": flops / (ms_kernel * 1e-3) / 1e12 / tf_peak, "traffic": None, "peak_kind": peak_kind + " (burst)",
                     "kernel": "tc_score_kernel", "algorithmic_flops_per_launch": flops, "ms_per_launch": ms_kernel},
        "e2e": {"value": Qn * e2e_steps / e2e_dt, "unit": "queries/s", "h2d_bytes_per_step": Qn * d * 4,
                "d2h_bytes_per_step": Qn * k * 12, "ms_per_step": e2e_dt / e2e_steps * 1e3,
                "api": "rag_docvqa_b200.sharded.search (pinned host questions in, (Q,k) scores + global ids out)"},
        "gpu_launches": args.steps * (5 if world == 1 else 6),
        "clocks": clocks.summary(),
    }
    if world == 1:
        # bf16 mode against the fp32 result (north_star): recall@k of the tensor-core path for the first 32 questions
        # against fp32 cosine scores of the same stored rows (un-rounded fp32 questions, plain torch matmul in fp32,
        # row chunks of 1 M).  Sharding does not change the hits (sharded == unsharded bit for bit, tests/test_tc_gpu.py),
        # so it is measured on the single-GPU run only.  Verification code after the timed region; never a bench value.
        try:
            n_q = min(32, Qn)
            allow = torch.backends.cuda.matmul.allow_tf32
            torch.backends.cuda.matmul.allow_tf32 = False
            _, got_idx = sharded.search(shard, q_dev, k)
            qs = q_dev[:n_q].float()
            qs = qs / qs.norm(dim=1, keepdim=True)
            best_v = torch.full((n_q, k), float("-inf"), device=dev)
            best_i = torch.zeros((n_q, k), dtype=torch.int64, device=dev)
            for a in range(0, hi - lo, 1 << 20):
                b = min(hi - lo, a + (1 << 20))
                r = rows[a:b].float()
                sc = (qs @ r.T) / r.norm(dim=1)
                v, i = sc.topk(min(k, b - a), dim=1)
                cat_v, cat_i = torch.cat([best_v, v], dim=1), torch.cat([best_i, i + a + lo], dim=1)
                best_v, pos = cat_v.topk(k, dim=1)
                best_i = torch.gather(cat_i, 1, pos)
                del r, sc
            torch.backends.cuda.matmul.allow_tf32 = allow
            got = got_idx[:n_q].to(torch.int64).cpu().numpy()
            ref_i = best_i.cpu().numpy()
            hits = sum(len(set(got[j].tolist()) & set(ref_i[j].tolist())) for j in range(n_q))
            line["recall_at_k_vs_fp32"] = {"value": hits / float(n_q * k), "questions": n_q, "k": k,
                                           "reference": "fp32 cosine (torch, TF32 off) of the fp32 questions against the stored bf16 rows"}
        except Exception as exc:                                   # reporting only: the bench line must still be printed
            line["recall_at_k_vs_fp32"] = {"value": None, "error": "%s: %s" % (type(exc).__name__, exc)}
    if rank == 0 and world == 1:
        # CPU: 1/64 row slice via torch.matmul + topk, scaled
        from oracle import ref_restated as R
        threads = os.cpu_count() or 1
        torch.set_num_threads(threads)
        n_cpu = max(1, N // 64)
        e_cpu = rows[:n_cpu].float().cpu()
        q_cpu = q_host.clone()

        def cpu_fn():
            return torch.topk(R.corpus_scores(e_cpu, q_cpu), k, dim=1)
        best, reps = time_cpu(cpu_fn, min(args.cpu_seconds, 10.0))
        line["cpu_baseline"] = {"value": Qn / (best * 64), "unit": "queries/s", "cores": threads, "kind": "port",
                                "sample": "oracle corpus_scores (torch matmul) + torch.topk on a 1/64 row slice (%d rows), "
                                          "time scaled x64, best of %d reps" % (n_cpu, reps)}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------
# visual path (BASELINE.json configs[3]): MaxSim late interaction of question tokens against page strips
# ------------------------------------------------------------------------------------------------
def run_visual(args):
    """C4: B documents x 50 strips x (2048 x 768) un-pooled encoder tokens, one (2048 x 768) question each
    (reference src/_modules.py:2191-2205 + src/utils.py:442-458, then torch.topk :2408).  A step = the MaxSim
    scores of one batch of B questions + the per-document top-k."""
    from rag_docvqa_b200 import functional as F
    from rag_docvqa_b200 import synth
    from rag_docvqa_b200.retriever import VisualRetriever
    rank, world, local = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    _, tf_peak, peak_kind = measured_peaks()
    B, strips, L, d, k = args.visual_docs, 50, 2048, 768, 5
    patches, q = synth.make_strip_batch(B, [strips] * B, L, d, synth_seed(4) + 1000 * rank, device=dev)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world > 1:
            t = torch.tensor([x], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return x

    vr_step = VisualRetriever({"chunk_num": k, "include_surroundings": 0, "chunk_mode": "horizontal", "device": str(dev)})

    def step(i):
        sims = vr_step._get_similarities(patches, q)            # the drop-in's own scoring loop (two side streams)
        return F.topk_segments(sims, k)

    # the dominant kernel alone: operands already normalised and split
    qs = [F.split_tf32(q[b], normalise=True) for b in range(B)]
    ps = [F.split_tf32(patches[b], normalise=True) for b in range(B)]
    from rag_docvqa_b200 import _lib
    tiles = (L + 127) // 128
    partial = torch.empty(strips * tiles, device=dev)
    out = torch.empty(strips, device=dev)
    stream = torch.cuda.current_stream(dev).cuda_stream

    def kernel_only(i):
        b = i % B
        _lib.check(_lib.lib.rdv_maxsim_tf32x3_tc(qs[b][0].data_ptr(), qs[b][1].data_ptr(), ps[b][0].data_ptr(), ps[b][1].data_ptr(),
                                                 strips, L, L, d, partial.data_ptr(), out.data_ptr(), stream))
    warmup = max(3, args.warmup)
    steps = max(1, min(args.steps, 20))
    for i in range(warmup):
        step(i)
        kernel_only(i)
    with ClockSampler(local) as clocks:
        ms_total = timed_loop(step, steps, barrier)
        ms_kernel = timed_loop(kernel_only, steps * B, barrier) / (steps * B)
    ms_per_step = max_over_ranks(ms_total) / steps
    ms_kernel = max_over_ranks(ms_kernel)
    del qs, ps
    flops = 2.0 * strips * L * L * d                          # fp32 contraction of one question
    ceiling = tf_peak / 2 / 3                                  # tf32 = half the bf16 rate; 3 tf32 products per fp32 product

    # e2e: the drop-in VisualRetriever.retrieve with pinned HOST token matrices, PIL pages in, crops + page ids out
    from PIL import Image
    pages = [[Image.new("RGB", (212, 275), (b * 7 % 255, g * 5 % 255, 0)) for g in range(strips)] for b in range(B)]
    flat = [np.arange(strips, dtype=np.int64) for _ in range(B)]
    mats = [[[[pages[b][g]]] for g in range(strips)] for b in range(B)]
    xyxy = [[[[0, 0, 212, 275]] for g in range(strips)] for b in range(B)]
    host_p = [x.cpu().pin_memory() for x in patches]
    host_q = q.cpu().pin_memory()
    vr = VisualRetriever({"chunk_num": k, "include_surroundings": 0, "chunk_mode": "horizontal", "device": str(dev)})

    def e2e_step():
        return vr.retrieve(host_p, host_q, flat, mats, xyxy, pages)
    e2e_step()
    barrier()
    e2e_steps = 3
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_dt = max_over_ranks(time.perf_counter() - t0)
    h2d = sum(x.numel() * 4 for x in host_p) + host_q.numel() * 4
    line = {
        "metric": METRIC, "value": B * world / (ms_per_step * 1e-3), "unit": "queries/s", "n_gpus": world, "steps": steps,
        "warmup": warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "tf32x3 (fp32 operands split exactly into two tf32 parts, fp32 accumulation)", "data": "synthetic",
        "config": {"workload": "C4: %d questions x %d strips x (%d x %d) tokens, MaxSim late interaction, top-k=%d" % (B, strips, L, d, k),
                   "step": "per document: F.normalize + tf32 hi/lo split of question and strips, tcgen05 kind::tf32 MaxSim, "
                           "strip sums; then one segmented top-k kernel for the batch",
                   "l2": "inputs larger than L2 (%.0f MB of strip tokens per document)" % (strips * L * d * 4 / 1e6),
                   "parallelism": "documents sharded across ranks (dp%d), no data-path collective" % world},
        "roofline": {"bound": "tensor", "achieved": flops / (ms_kernel * 1e-3) / 1e12, "peak": ceiling, "unit": "TFLOP/s",
                     "frac": flops / (ms_kernel * 1e-3) / 1e12 / ceiling, "traffic": None,
                     "peak_kind": peak_kind + " bf16 burst / 2 (tf32 rate) / 3 (products per fp32 product)",
                     "kernel": "maxsim_tf32x3_kernel", "algorithmic_flops_per_launch": flops, "ms_per_launch": ms_kernel},
        "e2e": {"value": B * world * e2e_steps / e2e_dt, "unit": "queries/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": B * (k + 1) * 4, "ms_per_step": e2e_dt / e2e_steps * 1e3,
                "api": "rag_docvqa_b200.retriever.VisualRetriever.retrieve (reference signature; pinned host token matrices and "
                       "PIL pages in; crops + page ids out)"},
        "gpu_launches": steps * (B * 5 + 1),
        "clocks": clocks.summary(),
    }
    if rank == 0 and world == 1:
        from oracle import ref_restated as R
        threads = os.cpu_count() or 1
        torch.set_num_threads(threads)
        if not args.no_extras:
            # a12, Pix2Struct half: the 5 retrieved strips of every document -> (2048, 770) flattened patches on the device
            from rag_docvqa_b200.pagestore import PageStore
            rng = np.random.RandomState(7)
            strip_pages = [[Image.fromarray(rng.randint(0, 256, (220, 850, 3)).astype(np.uint8), "RGB") for _ in range(k)] for _ in range(B)]
            pstore = PageStore.from_images(strip_pages, dev)
            crops = [[(g, 0, 0, 850, 220) for g in range(k)] for _ in range(B)]
            for _ in range(2):
                pstore.pack_pix2struct(crops)
            ms_p = timed_loop(lambda i: pstore.pack_pix2struct(crops), 10, torch.cuda.synchronize) / 10
            arrs = [np.asarray(im) for im in strip_pages[0]]
            best_p, _ = time_cpu(lambda: R.pix2struct_patches(arrs, 2048), 2.0, min_reps=2)
            line["extras"] = {"pix2struct_patches": {
                "what": "%d documents x %d strips of 850 x 220 -> (2048, 770) flattened patches + mask (plan on the host, "
                        "4 kernels)" % (B, k), "ms_per_batch": ms_p, "documents_per_s": B / ms_p * 1e3,
                "cpu_reference_ms_per_document": best_p * 1e3, "cpu_documents_per_s": 1.0 / best_p}}
        n_cpu = 5
        qc, pc = q[0:1].cpu(), patches[0][:n_cpu].cpu()
        best, reps = time_cpu(lambda: R.late_interaction(qc, pc), min(args.cpu_seconds, 10.0), min_reps=1)
        line["cpu_baseline"] = {"value": 1.0 / (best * strips / n_cpu), "unit": "queries/s", "cores": threads, "kind": "port",
                                "sample": "oracle late_interaction (torch CPU: F.normalize + bmm + max + sum) on %d of the %d strips "
                                          "of one question, time scaled x%d, best of %d reps" % (n_cpu, strips, strips // n_cpu, reps)}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def synth_seed(config_id):
    from rag_docvqa_b200 import synth
    return synth.SEED_BASE + config_id


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C2", choices=["C1", "C2", "C3", "C4", "C5"])
    ap.add_argument("--visual-docs", type=int, default=8)
    ap.add_argument("--corpus-rows", type=int, default=10_000_000)
    ap.add_argument("--corpus-queries", type=int, default=1024)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--skip-e2e", action="store_true", help="profiling runs only: leave out the host-input arm")
    ap.add_argument("--algo", type=int, default=0, help="0 auto, 1 LDG kernel, 2 TMA kernel")
    ap.add_argument("--lanes", type=int, default=8, choices=[1, 2, 3, 4, 8],
                    help="captured streams the steps alternate between (1 = one dependent chain)")
    args = ap.parse_args()
    if args.skip_e2e:
        args.no_extras = True
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "C5":
        run_corpus(args)
    elif args.workload == "C4":
        run_visual(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
