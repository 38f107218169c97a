"""Turns the ncu evidence of a round into the committed summaries under profiles/.

    python scripts/summarise_ncu.py r2

Inputs (gpurun_out/): <tag>*.ncu-rep captures and / or <tag>*_raw.csv files (`ncu -i x.ncu-rep --page raw --csv`, written on
the GPU box when the captures are too large to travel) and <tag>_launches_c2.csv (the launch list of the bench step).
Outputs (profiles/): <tag>_ncu_summary.md, <tag>_launches_c2.csv, and traffic.json (per-launch DRAM bytes of the kernels
bench.py quotes in roofline.traffic, with the capture file and the commit they were taken at)."""
import collections
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__cluster_size", "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum", "launch__shared_mem_per_block_dynamic",
    "launch__shared_mem_per_block_static",
]
SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def raw_rows(path):
    if path.endswith(".ncu-rep"):
        text = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    else:
        text = open(path).read()
    rows = [r for r in csv.reader(io.StringIO(text)) if len(r) > 5]
    return rows[0], rows[1], rows[2:]


def short(name):
    name = name.replace("void ", "").replace("rdv::", "").replace("(int)", "").replace("(bool)", "")
    return name.split("(")[0]


def num(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return None


def commit():
    try:
        return subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], stdout=subprocess.PIPE, text=True).stdout.strip()
    except Exception:
        return None


def main(tag):
    os.makedirs(OUT, exist_ok=True)
    gdir = os.path.join(ROOT, "gpurun_out")
    lines = ["# ncu summaries, round %s (B200, `--set full --clock-control none`; per-launch values; the LAST captured launch of "
             "every kernel instantiation)\n" % tag]
    traffic = {}
    files = sorted(f for f in os.listdir(gdir) if f.startswith(tag) and (f.endswith(".ncu-rep") or f.endswith("_raw.csv")))
    for fname in files:
        hdr, units, data = raw_rows(os.path.join(gdir, fname))
        kcol = hdr.index("Kernel Name")
        warm = "warm" in fname
        lines.append("## %s%s\n" % (fname, " (`--cache-control none`: caches NOT flushed between replays)" if warm else
                                   " (`--cache-control all`, the default: cold caches)"))
        last = collections.OrderedDict()
        for r in data:
            # the grid size tells launches of one instantiation at different workloads apart (C2 vs C3)
            grid = r[hdr.index("launch__grid_size")] if "launch__grid_size" in hdr else ""
            last[(short(r[kcol]), grid)] = r
        for (name, grid), r in last.items():
            lines.append("kernel: `%s`  (grid %s)\n" % (name[:120], grid))
            dram = sum((num(r[hdr.index(m)]) or 0) * SCALE.get(units[hdr.index(m)], 1)
                       for m in ("dram__bytes_read.sum", "dram__bytes_write.sum") if m in hdr)
            traffic["%s|grid %s|%s" % (name.split("<")[0], grid, "warm" if warm else "cold")] = {
                "kernel": name, "grid": grid, "dram_bytes_per_launch": dram,
                "duration_us": (num(r[hdr.index("gpu__time_duration.sum")]) or 0) *
                               ({"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(units[hdr.index("gpu__time_duration.sum")], 1.0)),
                "capture": fname, "cache_control": "none" if warm else "all"}
            lines.append("| metric | value | unit |\n|---|---|---|")
            for w in WANT:
                if w in hdr:
                    i = hdr.index(w)
                    lines.append("| %s | %s | %s |" % (w, r[i], units[i]))
            stalls = []
            for i, h in enumerate(hdr):
                if "pcsamp_warps_issue_stalled" in h and "not_issued" not in h:
                    v = num(r[i])
                    if v is not None:
                        stalls.append((v, h.replace("smsp__pcsamp_warps_issue_stalled_", "")))
            tot = sum(v for v, _ in stalls) or 1
            lines.append("\nstall samples: " + ", ".join("%s %.0f%%" % (n, 100 * v / tot) for v, n in sorted(stalls, reverse=True)[:6]) + "\n")
    launches = os.path.join(gdir, "%s_launches_c2.csv" % tag)
    if os.path.exists(launches):
        agg = collections.defaultdict(list)
        with open(launches) as f:
            rows = [r for r in csv.reader(f) if len(r) > 10]
        hdr = rows[0]
        unit = rows[1][-1] if len(rows) > 1 else "ns"
        for r in rows[1:]:
            v = num(r[-1])
            if v is None:
                continue
            agg[short(r[hdr.index("Kernel Name")])].append(v)
        lines.append("## launch list (`ncu --metrics gpu__time_duration.sum --clock-control none` over `bench.py --steps 20 --warmup 3 "
                     "--skip-e2e --no-legs`; cold-cache, serialised: compare SHARES, not absolutes)\n")
        lines.append("| kernel | launches | mean us | share of my-kernel time |\n|---|---|---|---|")
        mine = {k: v for k, v in agg.items() if not k.startswith("at::") and "elementwise" not in k and "nccl" not in k.lower()}
        total = sum(sum(v) for v in mine.values()) or 1
        for name, v in sorted(mine.items(), key=lambda kv: -sum(kv[1])):
            scale = 1e-3 if max(v) > 1000 else 1.0     # ns -> us
            lines.append("| %s | %d | %.2f | %.1f%% |" % (name[:100], len(v), sum(v) / len(v) * scale, 100 * sum(v) / total))
        with open(os.path.join(OUT, "%s_launches_c2.csv" % tag), "w") as f:
            f.write(open(launches).read())
    with open(os.path.join(OUT, "%s_ncu_summary.md" % tag), "w") as f:
        f.write("\n".join(lines) + "\n")
    tpath = os.path.join(OUT, "traffic.json")
    # a kernel's entry carries the commit its capture was taken at: entries whose capture did not change keep theirs
    head, old = commit(), {}
    if os.path.exists(tpath):
        prev = json.load(open(tpath))
        old = {k: dict(v, commit=v.get("commit", prev.get("commit"))) for k, v in prev.get("kernels", {}).items()}
    for k, v in traffic.items():
        o = old.get(k)
        same = o and o.get("capture") == v["capture"] and abs(o.get("duration_us", -1) - v["duration_us"]) < 1e-9
        v["commit"] = o["commit"] if same else head
    doc = {"_comment": "per-launch DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) from `ncu --set full` captures; "
                       "bench.py copies the entry of its dominant kernel into roofline.traffic together with `capture` and the entry's `commit` (the commit the capture was taken at; the top-level `commit` is the one the file was last regenerated at)",
           "commit": head, "round": tag, "kernels": traffic}
    with open(tpath, "w") as f:
        json.dump(doc, f, indent=1)
    print("\n".join(lines[:40]))
    print("... %d lines; %d kernels in traffic.json" % (len(lines), len(traffic)))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "r2")
