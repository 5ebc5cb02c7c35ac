"""Aggregate `ncu --page source --csv --print-source cuda,sass` per CUDA source line (all template instances together).
usage: ncu -i rep.ncu-rep --page source --csv --print-source cuda,sass | python tools/ncu_lines.py [kernel-substring] [top]
Prints, per kernel, the lines with the most stall samples: instructions, shared / global wavefront + sector counts
(actual vs ideal) and the dominant stall reasons."""
import csv, sys, collections

pat = sys.argv[1] if len(sys.argv) > 1 else ""
top = int(sys.argv[2]) if len(sys.argv) > 2 else 50
csv.field_size_limit(1 << 30)
COLS = ["# Samples", "Instructions Executed", "L1 Wavefronts Shared", "L1 Wavefronts Shared Ideal",
        "L2 Theoretical Sectors Global", "L2 Theoretical Sectors Global Ideal", "L1 Tag Requests Global",
        "stall_long_sb", "stall_barrier", "stall_no_inst", "stall_mio", "stall_short_sb", "stall_math", "stall_lg",
        "stall_wait", "stall_not_selected", "stall_selected", "stall_dispatch", "stall_branch_resolving"]
agg = collections.defaultdict(lambda: collections.defaultdict(float))   # (kernel, file, line) -> col -> sum
src = {}
fpath = func = None
hdr = None
line = None
for row in csv.reader(sys.stdin):
    if not row:
        continue
    if row[0] == "File Path":
        fpath = row[1].split("/")[-1]; hdr = None; continue
    if row[0] == "Function Name":
        func = row[1]; continue
    if row[0] == "Line No":
        hdr = row; idx = {h: i for i, h in enumerate(hdr) if h in COLS}; continue
    if hdr is None or func is None or (pat and pat not in func):
        continue
    if row[0]:
        line = int(row[0]); src[(fpath, line)] = row[1].strip()[:90]
    if len(row) < len(hdr) or not row[2]:
        continue
    key = (func.split("(")[0][:40], fpath, line)
    for h, i in idx.items():
        v = row[i]
        if v and v not in ("n/a",):
            try: agg[key][h] += float(v)
            except ValueError: pass

kernels = sorted({k[0] for k in agg})
for kn in kernels:
    items = [(k, v) for k, v in agg.items() if k[0] == kn]
    tot = collections.defaultdict(float)
    for _, v in items:
        for h, x in v.items(): tot[h] += x
    print("== %s : samples %d  inst %d  smem wavefronts %d (ideal %d)  global sectors %d (ideal %d)" % (
        kn, tot["# Samples"], tot["Instructions Executed"], tot["L1 Wavefronts Shared"], tot["L1 Wavefronts Shared Ideal"],
        tot["L2 Theoretical Sectors Global"], tot["L2 Theoretical Sectors Global Ideal"]))
    stl = sorted(((h, x) for h, x in tot.items() if h.startswith("stall_")), key=lambda t: -t[1])
    print("   stalls: " + "  ".join("%s %.1f%%" % (h[6:], 100 * x / max(1, tot["# Samples"])) for h, x in stl[:9]))
    items.sort(key=lambda kv: -kv[1]["# Samples"])
    print("   %-22s %6s %9s %9s %9s %9s %9s  %s" % ("file:line", "smp%", "inst", "smemWF", "ideal", "gSect", "ideal", "top stalls | source"))
    for (k, v) in items[:top]:
        st = sorted(((h[6:], x) for h, x in v.items() if h.startswith("stall_")), key=lambda t: -t[1])[:3]
        print("   %-22s %6.2f %9d %9d %9d %9d %9d  %s | %s" % (
            "%s:%d" % (k[1][:16], k[2]), 100 * v["# Samples"] / max(1, tot["# Samples"]), v["Instructions Executed"],
            v["L1 Wavefronts Shared"], v["L1 Wavefronts Shared Ideal"], v["L2 Theoretical Sectors Global"],
            v["L2 Theoretical Sectors Global Ideal"], " ".join("%s:%d" % (h, 100 * x / max(1, v["# Samples"])) for h, x in st),
            src.get((k[1], k[2]), "")))
