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


def ncu_traffic(kernel, grid):
    """Per-launch DRAM traffic of `kernel` at this grid size from the committed ncu capture (profiles/traffic.json, written
    by scripts/summarise_ncu.py with the capture file and the commit it was taken at), or None -- never a value of this run."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        with open(path) as f:
            doc = json.load(f)
        e = doc["kernels"].get("%s|grid %s|cold" % (kernel, grid))
        if e is None:
            return None
        return {"dram_bytes_per_launch": e["dram_bytes_per_launch"], "capture": e["capture"], "commit": e.get("commit", doc.get("commit")),
                "ncu_duration_us_cold_serialised": e.get("duration_us")}
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
    if args.workload == "C4p":
        return run_reference_pooled(args)
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
        "config": text_config(w),
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
        "config": corpus_config(N, d, Qn, k),
        "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": threads, "kind": "port",
                         "sample": "oracle corpus_scores + torch.topk on a 1/64 row slice (%d rows) per step, time x64" % n_cpu},
        "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


def run_reference_pooled(args):
    """CPU arm of C4p: the oracle's pooled-patch scores (the reference's mean_pooling + cosine, torch CPU ops) + torch.topk,
    one document per step."""
    from oracle import ref_restated as R
    from rag_docvqa_b200 import synth
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    strips, L, d, k = 50, 2048, 768, 5
    patches, q = synth.make_strip_batch(1, [strips], L, d, synth_seed(4))

    def step():
        sims, strip_scores, _ = R.pooled_patch_scores(patches, q)
        return torch.topk(sims[0], k), torch.topk(strip_scores[0], k)
    for _ in range(max(1, min(args.warmup, 2))):
        step()
    steps = max(1, min(args.steps, 10))
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": 1.0 / dt, "unit": "queries/s", "n_gpus": args.gpus, "steps": steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": pooled_config(args.visual_docs, strips, L, d, k),
        "cpu_baseline": {"value": 1.0 / dt, "unit": "queries/s", "cores": threads, "kind": "port",
                         "sample": "oracle pooled_patch_scores + torch.topk on one document per step"},
        "e2e": {"value": 1.0 / dt, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
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
        "config": visual_config(args.visual_docs, strips, L, d, 5),
        "cpu_baseline": {"value": 1.0 / dt, "unit": "queries/s", "cores": threads, "kind": "port",
                         "sample": "oracle late_interaction on %d of %d strips of one question per step, time x%d" % (n_cpu, strips, strips // n_cpu)},
        "e2e": {"value": 1.0 / dt, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
class Ctx:
    """One process per GPU: device, barrier, max / gather over ranks."""

    def __init__(self):
        self.rank, self.world, self.local = dist_env()
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device -- the B200 path has no CPU fallback")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=self.dev)
            self.dist = dist

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        torch.cuda.synchronize()

    def all_ranks(self, x):
        """[x of rank 0, x of rank 1, ...] on every rank."""
        if self.dist is None:
            return [float(x)]
        t = torch.tensor([float(x)], device=self.dev, dtype=torch.float64)
        out = torch.empty(self.world, device=self.dev, dtype=torch.float64)
        self.dist.all_gather_into_tensor(out, t)
        return [float(v) for v in out.tolist()]

    def maxr(self, x):
        return max(self.all_ranks(x))

    def close(self):
        if self.dist is not None:
            self.dist.destroy_process_group()


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


def rotation(step_bytes, lanes):
    """Distinct resident batches the steps rotate over: by the time a batch is read again at least 2 x L2 bytes of other
    batches have streamed through (plus the batches in flight in the other lanes); a multiple of `lanes` so that a
    batch's buffers are only ever touched by one lane."""
    R = int(np.ceil(2.0 * L2_BYTES / max(1, step_bytes))) + 1
    if lanes > 1:
        R += lanes
        R = (R + lanes - 1) // lanes * lanes
    return max(2, R)


def capture(dev, make_fn, n, lanes=1):
    """A CUDA graph of n steps: step i is make_fn(stream)(i), steps alternate over `lanes` captured streams."""
    g = torch.cuda.CUDAGraph()
    main = torch.cuda.Stream(dev)
    extra = [torch.cuda.Stream(dev) for _ in range(lanes - 1)]
    fns = [make_fn(st.cuda_stream) for st in [main] + extra]
    main.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.graph(g, stream=main):
        for st in extra:
            st.wait_stream(main)
        for i in range(n):
            fns[i % lanes](i)
        for st in extra:
            main.wait_stream(st)
    return g


def timed_replays(ctx, g, steps_in_graph, min_ms, min_reps):
    """The timed region: `reps` replays of a captured graph of `steps_in_graph` steps, one CUDA event between replays
    (on the launching stream), barrier + synchronize on both sides.  reps is the same on every rank and large enough
    for the region to last >= min_ms.  Returns per-step statistics in ms."""
    for _ in range(3):
        g.replay()
    ctx.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    one = max(e0.elapsed_time(e1) / 3, 1e-3)
    reps = int(min(4000, max(min_reps, np.ceil(min_ms / one))))
    reps = int(ctx.maxr(reps))
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    ctx.barrier()
    ev[0].record()
    for i in range(reps):
        g.replay()
        ev[i + 1].record()
    ctx.barrier()
    per = np.array([ev[i].elapsed_time(ev[i + 1]) for i in range(reps)]) / steps_in_graph
    return {"median": float(np.median(per)), "mean": float(per.mean()), "min": float(per.min()),
            "p90": float(np.percentile(per, 90)), "replays": reps, "region_ms": float(ev[0].elapsed_time(ev[reps]))}


def text_leg(ctx, args, wl, lanes, compact=False):
    """Per-document retrieval on resident inputs (C1 / C2 / C3): returns (result dict, live objects for the e2e leg).
    A step = the retrieval of ONE batch of B questions: with a pre-tokenised DocStore the ONE-launch kernel
    (rdv_retrieve_vt5_f32: score + per-document top-k + gather into the generator tensors); without one (C3) the
    product's score + top-k call (rdv_score_topk_f32, the algo rdv_score_plan picks)."""
    from rag_docvqa_b200 import functional as F
    from rag_docvqa_b200 import synth, _lib
    from rag_docvqa_b200.docstore import DocStore
    dev, rank = ctx.dev, ctx.rank
    w = synth.WORKLOADS[wl]
    hbm_peak, _, peak_kind = measured_peaks()
    with_lists = wl != "C3"
    # weak scaling: every rank gets documents of the SAME sizes (the seeded shape) with its own embedding values
    base_seed = synth.SEED_BASE + w.config_id
    rank_seed = base_seed + 100000 * rank
    host_batch = synth.make_text_batch(wl, with_lists=with_lists, share_image_pool=24, seed=base_seed)
    sizes = host_batch["sizes"]
    step_bytes = score_bytes(sizes, w.dim, w.k)
    R = rotation(step_bytes, lanes)
    K = args.steps if not compact else max(1, min(args.steps, 8))
    steps_in_graph = K * int(np.ceil(R / K)) if K < R else K          # every replay walks the whole rotation at least once
    batches = [synth.make_text_batch(wl, device=dev, seed=base_seed, emb_seed=rank_seed + 1000 * (r + 1)) for r in range(R)]
    tables = [F.build_doc_table(b["text_embeddings"], w.dim, dev, algo=args.algo) for b in batches]
    outs = [dict(sims=torch.empty(t.total_rows, dtype=torch.float32, device=dev),
                 idx=torch.empty((t.B, w.k), dtype=torch.int32, device=dev),
                 val=torch.empty((t.B, w.k), dtype=torch.float32, device=dev),
                 cnt=torch.empty((t.B,), dtype=torch.int32, device=dev)) for t in tables]
    lib = _lib.lib
    qs = [F._f32_contig_aligned(b["question_embeddings"]) for b in batches]
    plans, store, prompts, docstore_build_s = None, None, None, None
    if with_lists:
        table_w = synth.make_tokens_for_words(host_batch["words_text_chunks"], seed=3)
        t_store = time.perf_counter()
        store = DocStore.from_lists(host_batch["words_text_chunks"], host_batch["words_box_chunks"],
                                    host_batch["layout_labels_chunks"], host_batch["page_indices"],
                                    lambda wd: table_w.get(wd, [2]), dev, images=host_batch["images"])
        docstore_build_s = time.perf_counter() - t_store
        prompts = prompts_for(w.docs)
        plans = [store.prepare_gather(o["idx"], o["cnt"], prompts, max_len=512, sims=o["sims"], topk_val=o["val"],
                                      max_rows=t.max_rows) for o, t in zip(outs, tables)]
    # what the product's own calls launch (rdv_retrieve_plan); --one-launch forces the cluster kernels for a measurement
    one_launch = bool(plans) and (plans[0].prefers_one_launch(tables[0]) or
                                  (args.one_launch and plans[0].can_retrieve_in_one_launch(tables[0])))
    topk_cluster = plans is None and (tables[0].use_cluster(w.k) or (args.one_launch and tables[0].cluster_fits(w.k)))
    topk_algo = tables[0].algo
    torch.cuda.synchronize()

    def make_launchers(sp):
        """step / score-only closures launching on stream `sp` through the C ABI."""
        def score(i):
            t, o = tables[i % R], outs[i % R]
            rc = lib.rdv_score_f32(t.pointers()[0], t.total_tiles, t.tile_rows, t.algo, qs[i % R].data_ptr(), t.B,
                                   t.d, o["sims"].data_ptr(), sp)
            if rc:
                _lib.check(rc)

        if one_launch:
            def step(i):
                plans[i % R].launch_retrieve(tables[i % R], qs[i % R], outs[i % R]["sims"], sp)
        elif plans is not None:
            def step(i):
                score(i)
                plans[i % R].launch(sp)
        elif topk_cluster:
            def step(i):
                t, o = tables[i % R], outs[i % R]
                d_ctas, n_ctas, cl = t.cluster_pointers()
                rc = lib.rdv_score_topk_cluster_f32(d_ctas, n_ctas, cl, qs[i % R].data_ptr(), t.B, t.d, w.k, t.max_rows,
                                                    o["sims"].data_ptr(), o["idx"].data_ptr(), o["val"].data_ptr(),
                                                    o["cnt"].data_ptr(), sp)
                if rc:
                    _lib.check(rc)
        else:
            def step(i):
                t, o = tables[i % R], outs[i % R]
                rc = lib.rdv_score_topk_f32(t.pointers()[0], t.total_tiles, t.tile_rows, t.algo, t.pointers()[1],
                                            qs[i % R].data_ptr(), t.B, t.d, w.k, t.max_rows, o["sims"].data_ptr(),
                                            o["idx"].data_ptr(), o["val"].data_ptr(), o["cnt"].data_ptr(), sp)
                if rc:
                    _lib.check(rc)
        return step, score

    launches_per_step = 1 if (one_launch or topk_cluster) else 2
    stream = torch.cuda.current_stream(dev).cuda_stream
    launch_step, launch_score = make_launchers(stream)
    warmup = max(3, args.warmup)
    for i in range(max(warmup, R)):
        launch_step(i)
    torch.cuda.synchronize()

    g_chain = capture(dev, lambda sp: make_launchers(sp)[0], steps_in_graph, 1)
    g_lanes = capture(dev, lambda sp: make_launchers(sp)[0], steps_in_graph, lanes) if lanes >= 2 else None
    g_score = capture(dev, lambda sp: make_launchers(sp)[1], steps_in_graph, 1) if launches_per_step == 2 else None
    min_ms, min_reps = (args.min_ms, args.min_replays) if not compact else (args.min_ms / 2, 10)
    with ClockSampler(ctx.local) as clocks:
        st_lanes = timed_replays(ctx, g_lanes, steps_in_graph, min_ms, min_reps) if g_lanes is not None else None
        st_chain = timed_replays(ctx, g_chain, steps_in_graph, min_ms, min_reps)
        st_score = timed_replays(ctx, g_score, steps_in_graph, min_ms / 2, 10) if g_score is not None else None
        ms_plain = timed_loop(launch_step, steps_in_graph, ctx.barrier) / steps_in_graph
    headline = st_lanes if st_lanes is not None else st_chain
    per_rank = ctx.all_ranks(headline["median"])
    per_rank_chain = ctx.all_ranks(st_chain["median"])
    ms_per_step = max(per_rank)
    ms_chain = max(per_rank_chain)
    qps = w.docs * ctx.world / (ms_per_step * 1e-3)

    # algorithmic bytes of the gather half: records read per hit / per emitted token + the packed tensors written
    gather_bytes = 0
    if plans is not None:
        meta = plans[0].t["meta"].cpu().numpy()
        emitted = int(np.minimum(meta[0], 512).sum())
        gather_bytes = w.docs * w.k * 64 + emitted * 32 + w.docs * 512 * 48 + w.docs * w.k * (16 + 32 + 16)
    if launches_per_step == 1:
        kernel = ("retrieve_cluster_kernel<mode 2> (rdv_retrieve_vt5_f32: score + top-k + gather, a cluster per document)"
                  if one_launch else "retrieve_cluster_kernel<mode 1> (rdv_score_topk_cluster_f32: score + top-k)")
        kernel_bytes, kernel_ms = step_bytes + gather_bytes, ms_chain
        how = ("the step IS this one kernel: launches back to back in one dependent chain (CUDA graph, programmatic "
               "dependent launch), CUDA events around each replay, median over the replays, max over ranks")
    else:
        kernel = "score_tma_kernel" if topk_algo == _lib.SCORE_TMA else "score_ldg_kernel (rdv_score_f32)"
        kernel_bytes, kernel_ms = step_bytes, ctx.maxr(st_score["median"])
        how = ("the streaming kernel alone: launches back to back in one dependent chain (CUDA graph, programmatic dependent "
               "launch), CUDA events around each replay, median over the replays, max over ranks")
    achieved = kernel_bytes / (kernel_ms * 1e-3) / 1e9
    res = {
        "value": qps, "ms_per_step": ms_per_step, "steps_in_graph": steps_in_graph, "replays": headline["replays"],
        "timed_region_ms": headline["region_ms"], "rotation_batches": R, "rotation_MB": R * step_bytes / 1e6,
        "lanes": lanes if g_lanes is not None else 1, "launches_per_step": launches_per_step,
        "gpu_launches": launches_per_step * steps_in_graph * headline["replays"],
        "per_rank_ms_per_step": per_rank,
        "one_chain": {"ms_per_step": ms_chain, "queries_per_s": w.docs * ctx.world / (ms_chain * 1e-3),
                      "GBps": (step_bytes + gather_bytes) / (ms_chain * 1e-3) / 1e9,
                      "frac_hbm": (step_bytes + gather_bytes) / (ms_chain * 1e-3) / 1e9 / hbm_peak,
                      "per_rank_ms_per_step": per_rank_chain, "p90_ms": st_chain["p90"], "min_ms": st_chain["min"]},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                     "peak_kind": peak_kind, "kernel": kernel, "algorithmic_bytes_per_launch": kernel_bytes,
                     "ms_per_launch": kernel_ms, "how": how,
                     "pipelined_frac": (step_bytes + gather_bytes) / (ms_per_step * 1e-3) / 1e9 / hbm_peak,
                     "traffic": None},
        "stages": {"step_ms_graph_one_chain": ms_chain, "step_ms_graph_lanes": ms_per_step if g_lanes is not None else None,
                   "step_ms_plain_stream_launches": ctx.maxr(ms_plain),
                   "score_ms": ctx.maxr(st_score["median"]) if st_score is not None else None,
                   "step_bytes": step_bytes, "gather_bytes": gather_bytes,
                   "step_GBps": (step_bytes + gather_bytes) / (ms_per_step * 1e-3) / 1e9,
                   "step_frac_hbm": (step_bytes + gather_bytes) / (ms_per_step * 1e-3) / 1e9 / hbm_peak},
        "clocks": clocks.summary(),
    }
    if ctx.rank == 0:
        short_name = "retrieve_cluster_kernel" if launches_per_step == 1 else kernel.split(" ")[0]
        grid = tables[0].n_ctas if launches_per_step == 1 else tables[0].total_tiles
        traffic = ncu_traffic(short_name, grid)
        if traffic:
            res["roofline"]["traffic"] = traffic["dram_bytes_per_launch"]
            res["roofline"]["traffic_from"] = {k: traffic.get(k) for k in ("capture", "commit", "ncu_duration_us_cold_serialised")}
    live = dict(w=w, host_batch=host_batch, batches=batches, tables=tables, outs=outs, plans=plans, store=store,
                prompts=prompts, sizes=sizes, step_bytes=step_bytes, docstore_build_s=docstore_build_s, R=R)
    return res, live


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


def pool_leg(ctx, n_chunks, dim, hbm_peak):
    """Masked mean pooling (a1) of the token outputs of one batch's chunks: the embedder's tail, feeding the step."""
    from rag_docvqa_b200 import functional as F
    from rag_docvqa_b200 import synth
    embs, mask = synth.make_token_batch(n_chunks, dim, 7, device=ctx.dev, max_len=160)
    valid = int(mask.sum().item())
    n, L, d = embs.shape
    by = valid * d * 4 + n * L * 8 + n * d * 4
    for _ in range(3):
        F.mean_pooling(embs, mask)
    ms = timed_loop(lambda i: F.mean_pooling(embs, mask), 10, torch.cuda.synchronize) / 10
    return {"kernel": "mean_pool_kernel", "shape": [n, L, d], "valid_tokens": valid, "ms": ms, "algorithmic_bytes": by,
            "GBps": by / ms / 1e6, "frac_hbm": by / ms / 1e6 / hbm_peak}


def embed_leg(ctx, B, L, hbm_peak, with_cpu):
    """What follows the gather (section 8f rank 2, last clause): the generator's input embeddings of one batch's packed tensors,
    shared(ids) + SpatialEmbeddings(boxes) (src/VT5.py:194-204, src/_modules.py:70-86), t5-base sizes, one launch."""
    from rag_docvqa_b200.vt5_embed import SpatialEmbeddings, VT5InputEmbeddings
    D, V, n_pos = 768, 32128, 1024
    g = torch.Generator().manual_seed(synth_seed(9))
    w = {"x": torch.randn(n_pos, D, generator=g), "y": torch.randn(n_pos, D, generator=g), "g": 1 + 0.1 * torch.randn(D, generator=g),
         "b": 0.1 * torch.randn(D, generator=g), "W": torch.randn(D, D, generator=g) / D ** 0.5, "lb": 0.1 * torch.randn(D, generator=g),
         "shared": torch.randn(V, D, generator=g)}
    # packed tensors shaped like the gather's: a 20-token prompt on the full-page box, words of 1-3 tokens sharing a box, ~12 % padding
    ids = torch.randint(2, V, (B, L), generator=g)
    boxes = torch.zeros((B, L, 4), dtype=torch.int64)
    word = torch.randint(0, 1001, (B, L, 4), generator=g)
    new_word = torch.rand(B, L, generator=g) < 0.75
    # OCR geometry: words stand on text lines (~10 words per line share top and bottom, x runs left to right); the
    # uniformly random boxes of the second measurement are the worst case for the coordinate tables (no row is reused)
    lines = torch.zeros((B, L, 4), dtype=torch.int64)
    for b in range(B):
        n_words = L
        per_line = torch.randint(6, 15, (n_words,), generator=g)
        top = 20
        wi = 0
        while wi < n_words:
            k = int(per_line[wi])
            h = int(torch.randint(10, 25, (1,), generator=g))
            xs = torch.sort(torch.randint(30, 940, (k,), generator=g)).values
            ws = torch.randint(15, 60, (k,), generator=g)
            for j in range(min(k, n_words - wi)):
                lines[b, wi + j] = torch.tensor([int(xs[j]), top, min(1000, int(xs[j] + ws[j])), top + h])
            wi += k
            top = top + h + 6 if top + h + 40 < 1000 else 20
    boxes_random = torch.zeros((B, L, 4), dtype=torch.int64)
    for b in range(B):
        fill = int(L * (0.8 + 0.2 * torch.rand(1, generator=g).item()))
        idx = torch.cummax(torch.where(new_word[b], torch.arange(L), torch.zeros(L, dtype=torch.int64)), 0).values
        boxes_random[b] = word[b][idx]
        boxes[b] = lines[b][torch.cumsum(new_word[b].to(torch.int64), 0).clamp(max=L - 1)]
        for t in (boxes, boxes_random):
            t[b, :20] = torch.tensor([0, 0, 1000, 1000])
            t[b, fill:] = 0
        ids[b, fill:] = 0
    sp = SpatialEmbeddings(w["x"], w["y"], w["g"], w["b"], 1e-12, w["W"], w["lb"], device=ctx.dev)
    emb = VT5InputEmbeddings(sp, w["shared"])
    ids_d, boxes_d = ids.to(ctx.dev), boxes.to(ctx.dev)
    for _ in range(3):
        emb(ids_d, boxes_d)
    emb.check()
    reps = 20
    ms = timed_loop(lambda i: emb(ids_d, boxes_d), reps, torch.cuda.synchronize) / reps
    boxes_r = boxes_random.to(ctx.dev)
    for _ in range(3):
        emb(ids_d, boxes_r)
    ms_random = timed_loop(lambda i: emb(ids_d, boxes_r), reps, torch.cuda.synchronize) / reps
    n = B * L
    by = n * (D * 4 * 2 + 5 * 8)
    out = {"kernel": "vt5_embed_kernel", "shape": [B, L, D], "tokens_per_s": n / ms * 1e3, "ms": ms, "algorithmic_bytes": by,
           "GBps": by / ms / 1e6, "frac_hbm": by / ms / 1e6 / hbm_peak, "ms_uniformly_random_boxes": ms_random,
           "boxes": "words on text lines (6-14 words per line share top and bottom), 1-3 tokens per word, 20-token prompt, 0-20 % padding",
           "what": "input_embeds = shared(ids) + spatial_embedding(boxes) for one batch's packed tensors; algorithmic bytes = the row "
                   "written + the token's embedding row read + ids / boxes (the 32128 x 768 token table is 99 MB: part of it stays "
                   "in the 126 MB L2 between launches, as it does between batches in use; the coordinate tables are L2-resident by design)"}
    if with_cpu:
        from oracle import ref_restated as R
        threads = os.cpu_count() or 1
        torch.set_num_threads(threads)
        sb = max(1, min(B, 8))

        def cpu():
            spat = R.spatial_embeddings(boxes[:sb], w["x"], w["y"], w["g"], w["b"], 1e-12, w["W"], w["lb"])
            return R.vt5_input_embeds(ids[:sb], boxes[:sb], w["shared"], spat)
        best, _ = time_cpu(cpu, 2.0)
        out["cpu_tokens_per_s"] = sb * L / best
        out["cpu_sample"] = "oracle (the reference's torch CPU operators: 5 embedding lookups, LayerNorm, Linear) on %d of the %d rows, %d threads" % (sb, B, threads)
    del emb, sp
    return out


def e2e_text(ctx, args, live, lazy, cached=False):
    """The drop-in Retriever.retrieve with HOST inputs: pinned host embeddings + the reference's nested lists + PIL pages in,
    the 9-tuple out; H2D and D2H inside the timed region."""
    from rag_docvqa_b200 import functional as F
    from rag_docvqa_b200.retriever import Retriever
    w, dev, host_batch, batches, sizes = live["w"], ctx.dev, live["host_batch"], live["batches"], live["sizes"]
    host_sets = [([e.cpu().pin_memory() for e in b["text_embeddings"]], b["question_embeddings"].cpu().pin_memory())
                 for b in batches[:min(len(batches), 4)]]
    h2d = sum(e.numel() * 4 for e in host_sets[0][0]) + host_sets[0][1].numel() * 4
    cfg = {"compute_stats": False, "compute_stats_examples": False, "n_stats_examples": 0, "chunk_num": w.k,
           "device": str(dev)}
    if host_batch.get("words_text_chunks") is not None:
        lists = (host_batch["words_text_chunks"], host_batch["words_box_chunks"], host_batch["layout_labels_chunks"],
                 host_batch["images"], host_batch["page_indices"])
        retr = Retriever({**cfg, "retrieval_lazy_patches": bool(lazy), "retrieval_pause_gc": True,
                          "retrieval_embedding_cache_mb": 8192 if cached else 0})

        def step(i):
            emb_h, q_h = host_sets[i % len(host_sets)]
            return retr.retrieve(emb_h, q_h, *lists)
        d2h = w.docs * (w.k + 1) * 4 + sum(sizes) * 4
        api = ("rag_docvqa_b200.retriever.Retriever.retrieve (reference signature; pinned host embeddings, nested lists and "
               "PIL pages in; 9-tuple out; patches = %s; optional keys set: retrieval_pause_gc%s)" % (
                   "deferred crops (retrieval_lazy_patches)" if lazy else "eager PIL crops, as the reference",
                   ", retrieval_embedding_cache_mb (repeat questions about resident documents: no embedding crosses PCIe)"
                   if cached else ""))
        n_steps = max(5, min(args.steps, 40)) if lazy else 3
    else:
        cache = F.EmbeddingCache(64 << 30, dev) if cached else None

        def step(i):
            emb_h, q_h = host_sets[i % len(host_sets)]
            if cache is not None:
                table_h = F.build_doc_table(cache.resident(emb_h), w.dim, dev)
            else:
                table_h = F.upload_doc_table(emb_h, w.dim, dev)
            res = F.score_topk_table(table_h, q_h.to(dev, non_blocking=True), w.k)
            return res.topk_idx.cpu(), res.topk_cnt.cpu()
        d2h = w.docs * (w.k + 1) * 4
        api = ("rag_docvqa_b200.functional.%s + score_topk_table (pinned host embeddings in, top-k out)"
               % ("EmbeddingCache.resident + build_doc_table" if cached else "upload_doc_table"))
        n_steps = 5
    for i in range(max(3 if lazy else 1, len(host_sets) if cached else 0)):
        step(i)
    ctx.barrier()
    t0 = time.perf_counter()
    for i in range(n_steps):
        step(i)
    torch.cuda.synchronize()
    per_rank = ctx.all_ranks(time.perf_counter() - t0)
    dt = max(per_rank)
    return {"value": w.docs * ctx.world * n_steps / dt, "unit": "queries/s", "h2d_bytes_per_step": h2d,
            "d2h_bytes_per_step": d2h, "ms_per_step": dt / n_steps * 1e3, "steps": n_steps,
            "per_rank_ms_per_step": [t / n_steps * 1e3 for t in per_rank], "api": api}


CONFIG_L2 = "inputs larger than L2: every step streams a different resident batch (rotation > 2 x 126 MB)"
CONFIG_PAR = "documents sharded across ranks, no data-path collective"


def text_config(w):
    """The `config` object of a per-document workload -- identical in both arms (the driver compares them)."""
    return {"workload": workload_text(w), "l2": CONFIG_L2, "parallelism": CONFIG_PAR}


def corpus_config(N, d, Qn, k):
    return {"workload": "C5: %d chunks x %d-d bf16 row-sharded over the GPUs, %d questions, top-k=%d" % (N, d, Qn, k),
            "l2": "inputs larger than L2 (15.4 GB of corpus rows)",
            "parallelism": "corpus rows sharded across ranks, one NCCL all-gather of (Q, k) candidates per step"}


def run_ours(args):
    from rag_docvqa_b200 import synth
    ctx = Ctx()
    rank, world = ctx.rank, ctx.world
    w = synth.WORKLOADS[args.workload]
    hbm_peak, _, _ = measured_peaks()
    res, live = text_leg(ctx, args, args.workload, args.lanes)
    pipelined = res["lanes"] >= 2
    e2e = None
    if args.skip_e2e:
        e2e = {"value": None, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
               "api": "skipped (--skip-e2e: profiling run, keeps the kernel launch list to the device-resident steps)"}
    else:
        e2e = e2e_text(ctx, args, live, lazy=True)
    line = {
        "metric": METRIC, "value": res["value"], "unit": "queries/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(3, args.warmup), "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": text_config(w),
        "timing": {"what": "a step = the retrieval of one batch of %d questions = %s; %d steps are captured into one CUDA "
                           "graph and the timed region is %d replays of it (one CUDA event between replays, barrier + "
                           "synchronize on both sides); ms_per_step = median over the replays, max over ranks; %s"
                           % (w.docs,
                              "ONE launch (score + per-document top-k + gather into the packed VT5 inputs, max_source_length 512)"
                              if res["launches_per_step"] == 1 and live["plans"] is not None else
                              "ONE launch (score + per-document top-k)" if res["launches_per_step"] == 1 else
                              "streaming score kernel + select/gather kernel", res["steps_in_graph"], res["replays"],
                              "successive independent batches rotate over %d captured streams (a serving loop with %d batches in "
                              "flight); `one_chain` is the same steps as one dependent chain" % (res["lanes"], res["lanes"])
                              if pipelined else "one dependent chain"),
                   "steps_in_graph": res["steps_in_graph"], "replays": res["replays"], "timed_region_ms": res["timed_region_ms"],
                   "rotation_batches": res["rotation_batches"], "rotation_MB": res["rotation_MB"], "lanes": res["lanes"],
                   "per_rank_ms_per_step": res["per_rank_ms_per_step"]},
        "one_chain": res["one_chain"],
        "roofline": res["roofline"],
        "e2e": e2e,
        "gpu_launches": res["gpu_launches"],
        "clocks": res["clocks"],
        "stages": res["stages"],
    }
    if rank == 0 and world == 1:
        threads = os.cpu_count() or 1
        torch.set_num_threads(threads)
        cpu_batch = dict(live["host_batch"])
        cpu_batch["text_embeddings"] = [e.cpu() for e in live["batches"][0]["text_embeddings"]]
        cpu_batch["question_embeddings"] = live["batches"][0]["question_embeddings"].cpu()
        st_best, st_reps = time_cpu(cpu_score_topk_fn(cpu_batch, w.k), min(3.0, args.cpu_seconds))
        if live["plans"] is not None:
            best, reps = time_cpu(cpu_retrieve_fn(cpu_batch, w.k, crop=False), args.cpu_seconds)
            line["cpu_baseline"] = {
                "value": w.docs / best, "unit": "queries/s", "cores": threads, "kind": "port",
                "sample": "oracle retrieve() (score + torch.topk + Python list gather + crop rectangles, no PIL pixel "
                          "copy) on one full %s batch, best of %d reps" % (w.name, reps),
                "score_topk_only_queries_per_s": w.docs / st_best}
        else:
            line["cpu_baseline"] = {"value": w.docs / st_best, "unit": "queries/s", "cores": threads, "kind": "port",
                                    "sample": "oracle score+topk on one full %s batch, best of %d reps" % (w.name, st_reps)}
    if not args.skip_e2e and args.workload in ("C2", "C3"):
        # repeat questions about documents whose embeddings are already resident (MP-DocVQA: many questions per document):
        # the same call with the opt-in device cache -- NOT the headline e2e, whose every step crosses PCIe
        cached = e2e_text(ctx, args, live, lazy=True, cached=True)
        line["e2e_resident_cache"] = {"value": cached["value"], "unit": "queries/s", "ms_per_step": cached["ms_per_step"],
                                      "h2d_bytes_per_step": w.docs * w.dim * 4, "api": cached["api"]}
    if live["plans"] is not None and not args.skip_e2e and args.workload == "C2":
        # the reference's own output format: eager PIL crops in both arms (the crop is a host memcpy either way)
        eager = e2e_text(ctx, args, live, lazy=False)
        line["e2e_eager_pil_crops"] = {"value": eager["value"], "unit": "queries/s", "ms_per_step": eager["ms_per_step"]}
        if rank == 0 and world == 1:
            best_c, _ = time_cpu(cpu_retrieve_fn(cpu_batch, w.k, crop=True), 2.0, min_reps=1)
            line["e2e_eager_pil_crops"]["cpu_queries_per_s"] = w.docs / best_c
    extras = {}
    if args.extras and rank == 0 and world == 1 and live["plans"] is not None:
        extras.update(text_extras(ctx, args, live, cpu_batch))
    n_chunks = int(sum(live["sizes"]))
    step_bytes, gather_bytes = res["stages"]["step_bytes"], res["stages"]["gather_bytes"]
    del live, res
    torch.cuda.empty_cache()
    if not args.skip_e2e:
        # pooling (a1) at this batch's shape: with it the path is pool + score + top-k + gather (north_star's target)
        pool = pool_leg(ctx, n_chunks if args.workload != "C3" else 65536, w.dim, hbm_peak)
        tot_ms = pool["ms"] + line["one_chain"]["ms_per_step"]
        tot_by = pool["algorithmic_bytes"] + step_bytes + gather_bytes
        if args.workload != "C3":
            pool["pool_score_topk_gather"] = {"ms": tot_ms, "GBps": tot_by / tot_ms / 1e6, "frac_hbm": tot_by / tot_ms / 1e6 / hbm_peak,
                                              "what": "pooling of the batch's %d chunks followed by one step, one dependent chain" % n_chunks}
        line["pool"] = pool
        if args.workload == "C2":
            line["embed"] = embed_leg(ctx, w.docs, 512, hbm_peak, with_cpu=(rank == 0 and world == 1))
    if args.extras and rank == 0 and world == 1:
        extras.update(stage_extras(ctx.dev, hbm_peak, 20))
    if extras:
        line["extras"] = extras
    # ---- the other north_star configs, compact, as the LAST keys of the line -------------------------------------
    if args.workload == "C2" and not args.skip_e2e and not args.no_legs:
        c3, live3 = text_leg(ctx, args, "C3", 1, compact=True)
        line["c3"] = {"workload": workload_text(synth.WORKLOADS["C3"]), "queries_per_s": c3["value"], "ms_per_step": c3["ms_per_step"],
                      "kernel": c3["roofline"]["kernel"], "GBps": c3["roofline"]["achieved"], "frac_hbm": c3["roofline"]["frac"],
                      "step_frac_hbm": c3["one_chain"]["frac_hbm"], "launches_per_step": c3["launches_per_step"],
                      "replays": c3["replays"], "timed_region_ms": c3["timed_region_ms"],
                      "per_rank_ms_per_step": c3["per_rank_ms_per_step"], "scaling": "weak"}
        del c3, live3
        torch.cuda.empty_cache()
        c5 = corpus_leg(ctx, args, compact=True)
        line["corpus_c5"] = c5
        # the driver keeps the last 1500 characters of the line: both legs must fit there at N = 8
        line["c3"], line["corpus_c5"] = compact_leg(line["c3"]), compact_leg(line["corpus_c5"])
        for key in ("algorithmic_flops_per_launch", "traffic", "peak_kind"):
            line["corpus_c5"]["roofline"].pop(key, None)
        for key in ("per_rank_local_ms", "steps"):
            line["corpus_c5"].pop(key, None)
        line["corpus_c5"]["clocks"] = {k: line["corpus_c5"]["clocks"].get(k) for k in ("sm_mhz", "reasons")}
    if rank == 0:
        print(json.dumps(line))
    ctx.close()


def text_extras(ctx, args, live, cpu_batch):
    """Numbers for the widening rows on the C2 batch; rank 0, N = 1, --extras only."""
    from rag_docvqa_b200.retriever import Retriever
    w, dev, host_batch, batches, store, prompts = live["w"], ctx.dev, live["host_batch"], live["batches"], live["store"], live["prompts"]
    extras = {}
    lists = (host_batch["words_text_chunks"], host_batch["words_box_chunks"], host_batch["layout_labels_chunks"],
             host_batch["images"], host_batch["page_indices"])
    cfg = {"compute_stats": False, "compute_stats_examples": False, "n_stats_examples": 0, "chunk_num": w.k, "device": str(dev)}
    retr = Retriever({**cfg, "retrieval_lazy_patches": True})
    host_sets = [([e.cpu().pin_memory() for e in b["text_embeddings"]], b["question_embeddings"].cpu().pin_memory())
                 for b in batches[:2]]
    e2e_steps = 20
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
    extras["docstore_build_s_per_batch_of_documents"] = live["docstore_build_s"]
    extras["docstore_words"] = store.n_words
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
    pk, rs = retr.retrieve_packed(batches[0]["text_embeddings"], batches[0]["question_embeddings"], store, prompts)
    vplan = pstore.prepare_pack(pk.hit_page, pk.hit_rect, rs.topk_cnt)
    for _ in range(3):
        vplan.launch()
    ms_v = timed_loop(lambda i: vplan.launch(), 20, torch.cuda.synchronize) / 20
    area = int(((pk.hit_rect[..., 2] - pk.hit_rect[..., 0]).clamp(min=0) * (pk.hit_rect[..., 3] - pk.hit_rect[..., 1]).clamp(min=0)).sum().item())
    extras["visual_pack"] = {"ms": ms_v, "patch_pixels": area, "algorithmic_bytes": area * 3 + w.docs * 224 * 224 * 15,
                             "GBps": (area * 3 + w.docs * 224 * 224 * 15) / ms_v / 1e6,
                             "what": "64 documents x 5 crops -> grid canvas -> 224 x 224 bicubic (Pillow-exact) -> fp32 pixel_values"}
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
    from oracle import ref_restated as R_
    from rag_docvqa_b200.chunker import Chunker as _Chunker
    cw, cb, ci = synth_mod().make_chunker_batch(77, docs=16, max_pages=20, max_words=700, max_layouts=30, degenerate=False)
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
        "s": t_gpu, "cpu_oracle_s": t_cpu, "identical": got_chunks[0] == want_chunks[0] and got_chunks[2] == want_chunks[2]}
    return extras


def synth_mod():
    from rag_docvqa_b200 import synth
    return synth


# ------------------------------------------------------------------------------------------------
# corpus mode (BASELINE.json configs[4]): row-sharded bf16 corpus, tcgen05 scoring, NCCL all-gather + merge
# ------------------------------------------------------------------------------------------------
def corpus_leg(ctx, args, compact=False):
    """C5, strong scaling over the ranks: every rank owns N / world rows; a step answers Q questions against the whole
    corpus: bf16 cast of the questions, tcgen05 score + fused top-k on the local shard, local merge straight into the NCCL
    send buffer, all-gather of the (Q, k) candidates, final merge.  The step is captured once into a CUDA graph
    (sharded.CorpusSearcher) and replayed."""
    from rag_docvqa_b200 import sharded
    dev, rank, world = ctx.dev, ctx.rank, ctx.world
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
    searcher = sharded.CorpusSearcher(shard, Qn, k, group=None, graph=not args.no_graph)
    steps = max(1, min(args.steps, 20)) if not compact else max(1, min(args.steps, 10))
    warmup = max(3, args.warmup) if not compact else 3

    def step(i):
        return searcher.search(q_dev)
    for i in range(warmup):
        step(i)
    with ClockSampler(ctx.local) as clocks:
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
        ctx.barrier()
        ev[0].record()
        for i in range(steps):
            step(i)
            ev[i + 1].record()
        ctx.barrier()
        per = np.array([ev[i].elapsed_time(ev[i + 1]) for i in range(steps)])
        ms_kernel = timed_loop(lambda i: searcher.local_only(q_dev), steps, ctx.barrier) / steps
        ms_exchange = timed_loop(lambda i: searcher.exchange_only(), max(steps, 20), ctx.barrier) / max(steps, 20) if world > 1 else 0.0
    per_rank = ctx.all_ranks(float(np.median(per)))
    ms_per_step = max(per_rank)
    per_rank_kernel = ctx.all_ranks(ms_kernel)
    ms_kernel = max(per_rank_kernel)
    ms_exchange = ctx.maxr(ms_exchange)
    flops = 2.0 * Qn * (hi - lo) * d

    def e2e_step(i):
        val, idx = searcher.search(q_host.to(dev, non_blocking=True))
        return val.cpu(), idx.cpu()
    for i in range(2):
        e2e_step(i)
    ctx.barrier()
    e2e_steps = max(5, min(args.steps, 20)) if not compact else 5
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        e2e_step(i)
    torch.cuda.synchronize()
    e2e_dt = ctx.maxr(time.perf_counter() - t0)
    out = {
        "value": Qn / (ms_per_step * 1e-3), "ms_per_step": ms_per_step, "steps": steps, "per_rank_ms_per_step": per_rank,
        "roofline": {"bound": "tensor", "achieved": flops / (ms_kernel * 1e-3) / 1e12, "peak": tf_peak, "unit": "TFLOP/s",
                     "frac": flops / (ms_kernel * 1e-3) / 1e12 / tf_peak, "traffic": None, "peak_kind": peak_kind + " (burst)",
                     "kernel": "tc_score_kernel", "algorithmic_flops_per_launch": flops, "ms_per_launch": ms_kernel,
                     "how": ("local half alone, CUDA events, max over ranks" if compact else
                             "local half of the step alone (question cast + tc_score_kernel + local merge), CUDA events, max over ranks")},
        "local_ms": ms_kernel, "per_rank_local_ms": per_rank_kernel, "exchange_ms": ms_exchange,
        "tail_ms": ms_per_step - ms_kernel,
        "nccl_bytes_per_rank_per_step": Qn * k * 12 if world > 1 else 0,
        "e2e": {"value": Qn * e2e_steps / e2e_dt, "unit": "queries/s", "h2d_bytes_per_step": Qn * d * 4,
                "d2h_bytes_per_step": Qn * k * 12, "ms_per_step": e2e_dt / e2e_steps * 1e3,
                "api": ("sharded.CorpusSearcher.search, pinned host questions in" if compact else
                        "rag_docvqa_b200.sharded.CorpusSearcher.search (pinned host questions in, (Q,k) scores + global ids out)")},
        "graph": searcher.graphed, "clocks": clocks.summary(),
        "limiter": ("N=1: tc_score_kernel (tensor pipe / 1 kW power cap)" if world == 1 else
                    "tc_score_kernel %.3f ms alone (timed separately) against the %.3f ms step; all-gather + merges %.3f ms" % (ms_kernel, ms_per_step, ms_exchange)
                    if compact else
                    "tc_score_kernel %.3f ms of the %.3f ms step; the rest (%.3f ms) is the question cast, the two merges and "
                    "the all-gather (%.3f ms alone)" % (ms_kernel, ms_per_step, ms_per_step - ms_kernel, ms_exchange)),
    }
    # bf16 mode against the fp32 result (north_star): recall@k of the tensor-core hits of 32 questions against fp32 cosine scores
    # of the same stored rows (un-rounded fp32 questions, torch matmul in fp32, TF32 off), merged over the ranks' shards.
    # Verification code after the timed region; never a bench value.
    try:
        n_q = min(32, Qn)
        allow = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = False
        _, got_idx = searcher.search(q_dev)
        got = got_idx[:n_q].to(torch.int64).clone()
        qs = q_dev[:n_q].float()
        qs = qs / qs.norm(dim=1, keepdim=True)
        best_v = torch.full((n_q, k), float("-inf"), device=dev)
        best_i = torch.full((n_q, k), -1, dtype=torch.int64, device=dev)
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
        if world > 1:
            all_v = torch.empty((world, n_q, k), device=dev)
            all_i = torch.empty((world, n_q, k), dtype=torch.int64, device=dev)
            ctx.dist.all_gather_into_tensor(all_v, best_v.contiguous())
            ctx.dist.all_gather_into_tensor(all_i, best_i.contiguous())
            cat_v, cat_i = all_v.permute(1, 0, 2).reshape(n_q, -1), all_i.permute(1, 0, 2).reshape(n_q, -1)
            best_v, pos = cat_v.topk(k, dim=1)
            best_i = torch.gather(cat_i, 1, pos)
        got_np, ref_np = got.cpu().numpy(), best_i.cpu().numpy()
        hits = sum(len(set(got_np[j].tolist()) & set(ref_np[j].tolist())) for j in range(n_q))
        out["recall_at_k_vs_fp32"] = {"value": hits / float(n_q * k), "questions": n_q, "k": k,
                                      "reference": ("fp32 cosine of the stored rows" if compact else
                                                    "fp32 cosine (torch, TF32 off) of the fp32 questions against the stored bf16 rows")}
    except Exception as exc:                                   # reporting only: the bench line must still be printed
        out["recall_at_k_vs_fp32"] = {"value": None, "error": "%s: %s" % (type(exc).__name__, exc)}
    if rank == 0 and world == 1 and not compact:
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
        out["cpu_baseline"] = {"value": Qn / (best * 64), "unit": "queries/s", "cores": threads, "kind": "port",
                               "sample": "oracle corpus_scores (torch matmul) + torch.topk on a 1/64 row slice (%d rows), "
                                         "time scaled x64, best of %d reps" % (n_cpu, reps)}
    searcher.close()                      # before the process group is destroyed
    del searcher, shard, rows
    torch.cuda.empty_cache()
    return out


def run_corpus(args):
    ctx = Ctx()
    N, d, Qn, k = args.corpus_rows, 768, args.corpus_queries, 10
    c5 = corpus_leg(ctx, args)
    line = {
        "metric": METRIC, "value": c5["value"], "unit": "queries/s", "n_gpus": ctx.world, "steps": c5["steps"],
        "warmup": max(3, args.warmup), "ms_per_step": c5["ms_per_step"], "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": corpus_config(N, d, Qn, k),
        "timing": {"what": "a step = Q questions against the whole corpus: bf16 cast of the questions, tcgen05 score + fused top-k "
                           "on the local shard, local merge into the send buffer, NCCL all-gather of (Q,k) candidates, final merge"
                           + ("; captured once into a CUDA graph and replayed" if c5["graph"] else ""),
                   "per_rank_ms_per_step": c5["per_rank_ms_per_step"]},
        "roofline": c5["roofline"], "e2e": c5["e2e"],
        "gpu_launches": c5["steps"] * (5 if ctx.world == 1 else 6), "clocks": c5["clocks"],
        "stages": {k_: c5[k_] for k_ in ("local_ms", "per_rank_local_ms", "exchange_ms", "tail_ms", "nccl_bytes_per_rank_per_step",
                                         "limiter", "graph")},
        "recall_at_k_vs_fp32": c5.get("recall_at_k_vs_fp32"),
    }
    if "cpu_baseline" in c5:
        line["cpu_baseline"] = c5["cpu_baseline"]
    if ctx.rank == 0:
        print(json.dumps(line))
    ctx.close()


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
        "config": visual_config(B, strips, L, d, k),
        "timing": {"what": "a step = per document: F.normalize + tf32 hi/lo split of question and strips, tcgen05 kind::tf32 "
                           "MaxSim, strip sums; then one segmented top-k kernel for the batch"},
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
        if args.extras:
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


def pooled_config(B, strips, L, d, k):
    return {"workload": "C4p: %d questions x %d strips x %d patch vectors x %d-d, pooled question, cosine of every patch "
                        "vector, top-k=%d patches and strips" % (B, strips, L, d, k),
            "l2": "inputs larger than L2 (%.0f MB of patch vectors per document)" % (strips * L * d * 4 / 1e6),
            "parallelism": CONFIG_PAR}


def run_pooled(args):
    """C4p (north_star's wording of configs[3]): every patch vector of 50 strips (102 400 x 768 fp32 = 315 MB per
    document) scored against the mean-pooled question, top-k patches, strips ranked by their best patch.  A step = one batch
    of B questions through functional.pooled_patch_topk (pool, streaming score, top-k per strip, per-document selection)."""
    from rag_docvqa_b200 import functional as F
    from rag_docvqa_b200 import synth
    ctx = Ctx()
    dev, rank, world = ctx.dev, ctx.rank, ctx.world
    hbm_peak, _, peak_kind = measured_peaks()
    B, strips, L, d, k = args.visual_docs, 50, 2048, 768, 5
    patches, q = synth.make_strip_batch(B, [strips] * B, L, d, synth_seed(4) + 1000 * rank, device=dev)
    torch.cuda.synchronize()

    def step(i):
        return F.pooled_patch_topk(patches, q, k)
    flat = [p.reshape(-1, d) for p in patches]
    table = F.build_doc_table(flat, d, dev)
    q_pooled = step(0).question
    sims = torch.empty(table.total_rows, dtype=torch.float32, device=dev)
    warmup = max(3, args.warmup)
    steps = max(1, min(args.steps, 20))
    for i in range(warmup):
        step(i)
        F.score_table(table, q_pooled, out=sims)
    with ClockSampler(ctx.local) as clocks:
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
        ctx.barrier()
        ev[0].record()
        for i in range(steps):
            step(i)
            ev[i + 1].record()
        ctx.barrier()
        per = np.array([ev[i].elapsed_time(ev[i + 1]) for i in range(steps)])
        ms_kernel = timed_loop(lambda i: F.score_table(table, q_pooled, out=sims), steps, ctx.barrier) / steps
    per_rank = ctx.all_ranks(float(np.median(per)))
    ms_per_step = max(per_rank)
    ms_kernel = ctx.maxr(ms_kernel)
    n_rows = B * strips * L
    kernel_bytes = n_rows * d * 4 + B * d * 4 + n_rows * 4
    step_bytes_total = kernel_bytes + B * L * d * 4 + B * L * 8

    host_p = [x.cpu().pin_memory() for x in patches[:2]]
    host_q = q[:2].cpu().pin_memory()

    def e2e_step():
        res = F.pooled_patch_topk([x.to(dev, non_blocking=True) for x in host_p], host_q.to(dev, non_blocking=True), k)
        return res.patch_idx.cpu(), res.strip_idx.cpu()
    e2e_step()
    ctx.barrier()
    e2e_steps = 3
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_dt = ctx.maxr(time.perf_counter() - t0)
    line = {
        "metric": METRIC, "value": B * world / (ms_per_step * 1e-3), "unit": "queries/s", "n_gpus": world, "steps": steps,
        "warmup": warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": pooled_config(B, strips, L, d, k),
        "timing": {"what": "a step = one batch of %d questions through functional.pooled_patch_topk: mean pooling of the question "
                           "tokens, streaming cosine of every patch vector, top-k per strip, then per document the top-k of "
                           "the strips' candidates + strip scores + top-k strips (4 launches); CUDA events around every step, median, max over ranks" % B,
                   "per_rank_ms_per_step": per_rank},
        "roofline": {"bound": "hbm", "achieved": kernel_bytes / (ms_kernel * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                     "frac": kernel_bytes / (ms_kernel * 1e-3) / 1e9 / hbm_peak, "peak_kind": peak_kind, "traffic": None,
                     "kernel": "score_ldg_kernel (rdv_score_f32)", "algorithmic_bytes_per_launch": kernel_bytes,
                     "ms_per_launch": ms_kernel,
                     "how": "the streaming kernel alone, back-to-back launches, CUDA events, max over ranks",
                     "step_frac_hbm": step_bytes_total / (ms_per_step * 1e-3) / 1e9 / hbm_peak},
        "e2e": {"value": 2 * world * e2e_steps / e2e_dt, "unit": "queries/s", "h2d_bytes_per_step": sum(x.numel() * 4 for x in host_p) + host_q.numel() * 4,
                "d2h_bytes_per_step": 2 * 2 * k * 4, "ms_per_step": e2e_dt / e2e_steps * 1e3,
                "api": "rag_docvqa_b200.functional.pooled_patch_topk (pinned host token matrices of 2 documents in, hit indices out)"},
        "gpu_launches": steps * 4, "clocks": clocks.summary(),
    }
    if rank == 0 and world == 1:
        from oracle import ref_restated as R
        threads = os.cpu_count() or 1
        torch.set_num_threads(threads)
        pc, qc = [patches[0].cpu()], q[0:1].cpu()

        def cpu_fn():
            sims_c, strips_c, _ = R.pooled_patch_scores(pc, qc)
            return torch.topk(sims_c[0], k), torch.topk(strips_c[0], k)
        best, reps = time_cpu(cpu_fn, min(args.cpu_seconds, 10.0), min_reps=1)
        line["cpu_baseline"] = {"value": 1.0 / best, "unit": "queries/s", "cores": threads, "kind": "port",
                                "sample": "oracle pooled_patch_scores (torch CPU: mean pooling + cosine of 102 400 patch vectors + "
                                          "strip max) + torch.topk on ONE document, best of %d reps" % reps}
    if rank == 0:
        print(json.dumps(line))
    ctx.close()


def visual_config(B, strips, L, d, k):
    return {"workload": "C4: %d questions x %d strips x (%d x %d) tokens, MaxSim late interaction, top-k=%d" % (B, strips, L, d, k),
            "l2": "inputs larger than L2 (%.0f MB of strip tokens per document)" % (strips * L * d * 4 / 1e6),
            "parallelism": CONFIG_PAR}


def compact_leg(obj):
    """Floats to 5 significant digits, recursively (a leg of the default line, not a headline number)."""
    if isinstance(obj, float):
        return float("%.5g" % obj)
    if isinstance(obj, dict):
        return {k: compact_leg(v) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return [compact_leg(v) for v in obj]
    return obj


def synth_seed(config_id):
    from rag_docvqa_b200 import synth
    return synth.SEED_BASE + config_id


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C2", choices=["C1", "C2", "C3", "C4", "C4p", "C5"])
    ap.add_argument("--visual-docs", type=int, default=8)
    ap.add_argument("--corpus-rows", type=int, default=10_000_000)
    ap.add_argument("--corpus-queries", type=int, default=1024)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--extras", action="store_true", help="numbers for the widening rows (SURVEY 8f) and the other kernels; N=1 only")
    ap.add_argument("--no-extras", action="store_true", help="(default; kept for old command lines)")
    ap.add_argument("--no-legs", action="store_true", help="leave out the compact C3 / C5 legs of the default line")
    ap.add_argument("--one-launch", action="store_true", help="force the one-launch cluster kernels (measurement; the plan prefers two launches)")
    ap.add_argument("--no-graph", action="store_true", help="C5: eager launches instead of the captured step")
    ap.add_argument("--min-ms", type=float, default=25.0, help="shortest timed region (ms) of the replayed graph")
    ap.add_argument("--min-replays", type=int, default=50)
    ap.add_argument("--skip-e2e", action="store_true", help="profiling runs only: leave out the host-input arm")
    ap.add_argument("--algo", type=int, default=0, help="0 auto, 1 LDG kernel, 2 TMA kernel")
    ap.add_argument("--lanes", type=int, default=8, choices=[1, 2, 3, 4, 8],
                    help="captured streams the steps alternate between (1 = one dependent chain)")
    args = ap.parse_args()
    if args.skip_e2e:
        args.extras = False
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "C5":
        run_corpus(args)
    elif args.workload == "C4":
        run_visual(args)
    elif args.workload == "C4p":
        run_pooled(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
