"""Split the SASS of each kernel in an ncu source-page CSV at its BAR.SYNC instructions and print, per segment,
the share of stall samples, the instruction count and the executed warp instructions; then the hottest instructions."""
import csv, sys
path = sys.argv[1]; want = sys.argv[2] if len(sys.argv) > 2 else ''
rows = list(csv.reader(open(path)))
i = 0
while i < len(rows):
    if rows[i] and rows[i][0] == 'Kernel Name':
        name = rows[i][1]; hdr = rows[i + 1]; j = i + 2
        body = []
        while j < len(rows) and not (rows[j] and rows[j][0] == 'Kernel Name'):
            if len(rows[j]) == len(hdr): body.append(rows[j])
            j += 1
        i = j
        if want not in name: continue
        cs, ce, cx = hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
        tot = sum(int(r[ce] or 0) for r in body) or 1
        print("=====", name[:60], "samples", tot, "sass", len(body))
        seg = []; cur = [0, 0, 0, 0]
        for n, r in enumerate(body):
            cur[0] += int(r[ce] or 0); cur[1] += 1; cur[2] += int(r[cx] or 0)
            if 'BAR.SYNC' in r[cs] or 'EXIT' in r[cs]:
                seg.append((cur, n)); cur = [0, 0, 0, 0]
        for k, (c, n) in enumerate(seg):
            if c[0] * 200 > tot or c[2] > 0:
                print("  seg %2d ends@%5d: samples %5.1f%%  sass %5d  warp-inst %9d" % (k, n, 100.0 * c[0] / tot, c[1], c[2]))
        hot = sorted(body, key=lambda r: -int(r[ce] or 0))[:14]
        for r in hot:
            print("     %5.2f%%  %s" % (100.0 * int(r[ce] or 0) / tot, r[cs].strip()[:90]))
    else:
        i += 1
