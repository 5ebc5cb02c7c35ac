"""profiles/r2_launches.csv (ncu launch list of one bench step) -> profiles/r2_dram_traffic.json (read by bench.py for roofline.traffic).
usage: python tools/dram_traffic.py profiles/r2_launches.csv profiles/r2_dram_traffic.json"""
import csv, json, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
hdr = rows[hi]; kn = hdr.index('Kernel Name'); mn = hdr.index('Metric Name'); mv = hdr.index('Metric Value'); idc = hdr.index('ID')
d = {}
for r in rows[hi + 1:]:
    if len(r) > mv:
        d.setdefault((int(r[idc]), r[kn]), {})[r[mn]] = float(r[mv].replace(',', '') or 0)
names = {"slice_fft_fwd": "slice_fft_fwd", "bins_fwd": "bins_fwd", "bins_inv": "bins_inv", "slice_fft_inv": "slice_fft_inv"}
out = {k: {"launches_per_step": 0, "dram_bytes_per_step": 0.0, "ncu_time_us_per_step": 0.0, "warp_instructions_per_step": 0.0} for k in names}
for (i, k), m in d.items():
    for key in names:
        if key in k:
            o = out[key]
            o["launches_per_step"] += 1
            o["dram_bytes_per_step"] += m.get("dram__bytes_read.sum", 0) + m.get("dram__bytes_write.sum", 0)
            t = m.get("gpu__time_duration.sum", 0)
            o["ncu_time_us_per_step"] += t / 1e3 if t > 1e4 else t      # ns or us depending on the ncu unit setting
            o["warp_instructions_per_step"] += m.get("smsp__inst_executed.sum", 0)
            break
UA, US = 2368, 9472   # units of one bench step (batch of 8 mixtures): analysis / synthesis
tot = sum(o["ncu_time_us_per_step"] for o in out.values())
res = {"source": "profiles/r2_launches.csv: ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,... --clock-control none, tools/prof_step.py --batch 8 --steps 1 (one bench step: %d analysis + %d synthesis units)" % (UA, US),
       "analysis_dram_bytes_per_unit": (out["slice_fft_fwd"]["dram_bytes_per_step"] + out["bins_fwd"]["dram_bytes_per_step"]) / UA,
       "synthesis_dram_bytes_per_unit": (out["bins_inv"]["dram_bytes_per_step"] + out["slice_fft_inv"]["dram_bytes_per_step"]) / US,
       "ncu_time_share": {k: round(o["ncu_time_us_per_step"] / tot, 3) for k, o in out.items()},
       "kernels": out}
json.dump(res, open(sys.argv[2], "w"), indent=1)
print(json.dumps({k: res[k] for k in ("analysis_dram_bytes_per_unit", "synthesis_dram_bytes_per_unit", "ncu_time_share")}))
