"""Turns gpurun_out/*.ncu-rep + the launch list into the committed summaries under profiles/."""
import csv, io, json, os, subprocess, sys, collections

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum", "launch__shared_mem_per_block_dynamic",
]

def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]

TRAFFIC = {}


def main(round_tag):
    os.makedirs(OUT, exist_ok=True)
    lines = ["# ncu summaries, round %s (B200, `--set full --clock-control none`; per-launch values)\n" % round_tag]
    for rep in sorted(f for f in os.listdir(os.path.join(ROOT, "gpurun_out")) if f.startswith(round_tag) and f.endswith(".ncu-rep")):
        hdr, units, data = raw(os.path.join(ROOT, "gpurun_out", rep))
        kcol = hdr.index("Kernel Name")
        lines.append("## %s\n" % rep)
        for r in data:
            lines.append("kernel: `%s`\n" % r[kcol][:110])
            TRAFFIC[r[kcol].split("(")[0].replace("void ", "").split("<")[0].replace("rdv::", "")] = sum(
                float(r[hdr.index(m)].replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[units[hdr.index(m)]]
                for m in ("dram__bytes_read.sum", "dram__bytes_write.sum") if m in hdr)
            lines.append("| metric | value | unit |\n|---|---|---|")
            for w in WANT:
                if w in hdr:
                    i = hdr.index(w)
                    lines.append("| %s | %s | %s |" % (w, r[i], units[i]))
            stalls = []
            for i, h in enumerate(hdr):
                if "pcsamp_warps_issue_stalled" in h and "not_issued" not in h:
                    try:
                        stalls.append((float(r[i].replace(",", "")), h.replace("smsp__pcsamp_warps_issue_stalled_", "")))
                    except ValueError:
                        pass
            tot = sum(v for v, _ in stalls) or 1
            lines.append("\nstall samples: " + ", ".join("%s %.0f%%" % (n, 100 * v / tot) for v, n in sorted(stalls, reverse=True)[:6]) + "\n")
    launches = os.path.join(ROOT, "gpurun_out", "%s_launches_c2.csv" % round_tag)
    if os.path.exists(launches):
        agg = collections.defaultdict(list)
        with open(launches) as f:
            rows = [r for r in csv.reader(f) if len(r) > 10]
        hdr = rows[0]
        for r in rows[1:]:
            name = r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "").replace("rdv::", "")
            agg[name].append(float(r[-1].replace(",", "")))
        lines.append("## launch list (`--metrics gpu__time_duration.sum`, bench.py --steps 20 --warmup 3 --skip-e2e; cold-cache, serialised: compare SHARES)\n")
        lines.append("| kernel | launches | mean us | share of my-kernel time |\n|---|---|---|---|")
        total = sum(sum(v) for v in agg.values())
        for name, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            unit_scale = 1e-3 if max(v) > 1000 else 1.0     # ns -> us
            lines.append("| %s | %d | %.2f | %.1f%% |" % (name, len(v), sum(v) / len(v) * unit_scale, 100 * sum(v) / total))
        with open(os.path.join(OUT, "%s_launches_c2.csv" % round_tag), "w") as f:
            f.write(open(launches).read())
    with open(os.path.join(OUT, "%s_ncu_summary.md" % round_tag), "w") as f:
        f.write("\n".join(lines) + "\n")
    print("\n".join(lines))
    print("per-launch DRAM traffic (bytes):", json.dumps(TRAFFIC))

if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "r1")
