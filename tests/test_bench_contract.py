"""CPU-side checks of bench.py's contract: the reference arm runs without a GPU and prints the agreed line, the product arm
refuses to run without one (no CPU fallback), and the compact C3 / C5 legs of the default line stay inside the last 1500
characters, where the driver looks for them."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BENCH = os.path.join(ROOT, "bench.py")


def run(*args):
    return subprocess.run([sys.executable, BENCH, *args], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, cwd=ROOT,
                          env={**os.environ, "CUDA_VISIBLE_DEVICES": ""}, timeout=600)


def test_reference_arm_prints_the_contract_line_on_cpu():
    res = run("--impl", "reference", "--workload", "C1", "--steps", "2", "--warmup", "1")
    assert res.returncode == 0, res.stderr[-2000:]
    line = json.loads(res.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "retrieval_queries_per_sec" and line["unit"] == "queries/s"
    assert line["n_gpus"] == 1 and line["steps"] == 2 and line["warmup"] == 1 and line["higher_is_better"] is True
    assert line["value"] > 0 and line["ms_per_step"] > 0
    assert line["config"]["workload"].startswith("C1:")
    base = line["cpu_baseline"]
    assert base["kind"] in ("port", "reference") and base["cores"] >= 1 and base["sample"] and base["value"] == line["value"]
    e2e = line["e2e"]
    assert e2e["value"] == line["value"] and e2e["unit"] == line["unit"]
    assert e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0


def test_product_arm_fails_loudly_without_a_gpu():
    res = run("--workload", "C1", "--steps", "2", "--warmup", "1")
    assert res.returncode != 0
    assert "no CPU fallback" in (res.stderr + res.stdout)
    assert not res.stdout.strip().startswith("{")              # no bench line from a machine that cannot run the kernels


def test_compact_legs_fit_the_tail_of_the_line():
    sys.path.insert(0, ROOT)
    import bench
    assert bench.compact_leg({"a": 1.23456789, "b": [0.000123456789, 3], "c": "x", "d": {"e": 1e9 / 3}}) == {
        "a": 1.2346, "b": [0.00012346, 3], "c": "x", "d": {"e": 333330000.0}}
    for name in ("r2_bench_c2_n1.json", "r2_bench_c2_n8.json"):      # lines the builder measured on B200 boxes
        path = os.path.join(ROOT, "profiles", name)
        text = open(path).read().strip().splitlines()[-1]
        line = json.loads(text)
        keys = list(line)
        assert keys[-2:] == ["c3", "corpus_c5"], keys[-4:]
        tail = text[-1500:]
        assert '"c3": {' in tail and '"corpus_c5": {' in tail, "%s: the legs need %d characters" % (name, len(text) - text.index('"c3": {'))
        assert line["corpus_c5"]["recall_at_k_vs_fp32"]["value"] >= 0.9
        assert line["n_gpus"] == len(line["corpus_c5"]["per_rank_ms_per_step"])
