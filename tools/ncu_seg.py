"""Per-segment (BAR.SYNC delimited) stall-reason shares from an ncu source-page CSV; optional dump of one segment."""
import csv, sys
path, want = sys.argv[1], sys.argv[2]
dump = int(sys.argv[3]) if len(sys.argv) > 3 else -1
rows = list(csv.reader(open(path)))
i = [n for n, r in enumerate(rows) if r and r[0] == 'Kernel Name' and want in r[1]][0]
hdr = rows[i + 1]
body = []
for r in rows[i + 2:]:
    if r and r[0] == 'Kernel Name': break
    if len(r) == len(hdr): body.append(r)
cs, ce, cx = hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
st = [(n, h) for n, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
tot = sum(int(r[ce] or 0) for r in body) or 1
segs = []; cur = []
for r in body:
    cur.append(r)
    if 'BAR.SYNC' in r[cs] or 'EXIT' in r[cs]: segs.append(cur); cur = []
for k, sg in enumerate(segs):
    s = sum(int(r[ce] or 0) for r in sg)
    if s * 100 < tot: continue
    acc = {h: sum(int(r[n] or 0) for r in sg) for n, h in st}
    top = sorted(acc.items(), key=lambda t: -t[1])[:5]
    print("seg %2d  %5.1f%%  sass %4d  winst/cta %6d | " % (k, 100.0 * s / tot, len(sg), sum(int(r[cx] or 0) for r in sg) // 1184) + "  ".join("%s %.0f%%" % (h[6:], 100.0 * v / max(s, 1)) for h, v in top))
    if k == dump:
        for r in sg:
            if int(r[ce] or 0) * 400 > s:
                acc = sorted(((int(r[n] or 0), h[6:]) for n, h in st), reverse=True)[:2]
                print("      %5.1f%% %-70s %s" % (100.0 * int(r[ce]) / s, r[cs].strip()[:70], acc))
