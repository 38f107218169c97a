"""Top CUDA source lines by stall samples from an .ncu-rep (needs -lineinfo + --import-source on).
usage: python scripts/ncu_source_hot.py <rep> [top_n]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
sections, cur, meta = [], None, []
for r in rows:
    if len(r) > 4 and r[0] == "Line No":
        cur = {"hdr": r, "rows": [], "meta": meta}
        meta = []
        sections.append(cur)
    elif len(r) == 2 and r[0] in ("File Path", "Function Name", "Kernel Name"):
        meta.append(r[1][:100])
        cur = None
    elif cur is not None and len(r) == len(cur["hdr"]) and r[2] == "-":     # a CUDA source line (aggregate of its SASS)
        cur["rows"].append(r)
for s in sections:
    h = s["hdr"]
    ci = h.index("# Samples")
    stall_cols = [i for i, n in enumerate(h) if n.startswith("stall_") and "Not Issued" not in n]
    tot = sum(int(r[ci] or 0) for r in s["rows"])
    if tot == 0:
        continue
    print("== %s: %d samples" % (" | ".join(s["meta"]), tot))
    for r in sorted(s["rows"], key=lambda r: -int(r[ci] or 0))[:top]:
        n = int(r[ci] or 0)
        if n == 0:
            break
        st = sorted(((int(r[i] or 0), h[i][6:]) for i in stall_cols), reverse=True)[:3]
        print("%5.1f%% L%-4s %-100s %s" % (100.0 * n / tot, r[0], r[1].strip()[:100], " ".join("%s:%d" % (b, a) for a, b in st if a)))
