"""Summarise an ncu --csv launch list (gpu__time_duration etc.) per launch."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
hdr = rows[hi]; kn = hdr.index('Kernel Name'); mn = hdr.index('Metric Name'); mv = hdr.index('Metric Value'); idc = hdr.index('ID')
d = {}
for r in rows[hi + 1:]:
    if len(r) <= mv: continue
    d.setdefault((int(r[idc]), r[kn][:30]), {})[r[mn]] = r[mv].replace(',', '')
for (i, k), m in sorted(d.items()):
    f = lambda n: float(m.get(n, '0') or 0)
    print("%3d %-30s grid %5d  %8.1f us  %6.1f Minst  issue %4.1f%%  rd %6.1f MB  wr %6.1f MB" % (i, k, f('launch__grid_size'), f('gpu__time_duration.sum') / 1e3, f('smsp__inst_executed.sum') / 1e6, f('smsp__issue_active.avg.pct_of_peak_sustained_active'), f('dram__bytes_read.sum') / 1e6, f('dram__bytes_write.sum') / 1e6))
