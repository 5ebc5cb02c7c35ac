"""Print pc-sampling stall shares per kernel from an ncu raw-page CSV, and the hottest source lines."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
for r in rows[2:]:
    name = r[hdr.index('Kernel Name')][:40]
    print("=====", name, r[hdr.index('gpu__time_duration.sum')], "us")
    items = []
    for i, h in enumerate(hdr):
        if h.startswith('smsp__pcsamp_warps_issue_stalled_') and not h.endswith('_not_issued'):
            try: items.append((float(r[i]), h.replace('smsp__pcsamp_warps_issue_stalled_', '')))
            except ValueError: pass
    tot = sum(v for v, _ in items) or 1.0
    items.sort(reverse=True)
    print("   " + "  ".join("%s %.1f%%" % (h, 100 * v / tot) for v, h in items[:10]))
