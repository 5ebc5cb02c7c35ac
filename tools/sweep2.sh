#!/bin/bash
# usage (under gpurun): tools/sweep2.sh <variant> ...   -> one line per variant library (build/variants/<v>/libslicq.so; 'base' = in-tree)
for v in "$@"; do
  lib=$PWD/build/variants/$v/libslicq.so
  [ "$v" = base ] && lib=$PWD/xumx_slicq_b200/libslicq.so
  out=gpurun_out/sw_$v.json
  SLICQ_B200_LIB=$lib python bench.py --steps 8 --warmup 3 --no-cpu-baseline > $out 2>gpurun_out/sw_err.log || { echo "$v FAILED"; tail -3 gpurun_out/sw_err.log; continue; }
  python - "$v" "$out" <<'PY'
import json,sys
d=json.load(open(sys.argv[2])); k=d["roofline"]["kernels"]
print("%-14s ms/step %7.3f frac %.4f | "%(sys.argv[1], d["ms_per_step"], d["roofline"]["frac"]) + " ".join("%s=%.3f"%(n.replace("slice_fft","sf"),v["ms_per_step"]) for n,v in k.items()), "err %.1e"%d["max_abs_err_target0"])
PY
done
