#!/bin/bash
# usage: tools/sweep.sh "<variant>[,ENV=VAL...]" ...   -> one compact line per configuration (bench.py, B from $SWEEP_BATCH)
B=${SWEEP_BATCH:-8}
for spec in "$@"; do
  IFS=',' read -ra parts <<< "$spec"
  v=${parts[0]}
  envs=("${parts[@]:1}")
  out=gpurun_out/sweep_${spec//[^A-Za-z0-9_=]/_}.json
  env SLICQ_B200_LIB=$PWD/build/variants/$v/libslicq.so "${envs[@]}" python bench.py --steps 8 --warmup 3 --batch $B --no-cpu-baseline > $out 2>gpurun_out/sweep_err.log || { echo "$spec FAILED"; tail -3 gpurun_out/sweep_err.log; continue; }
  python - "$spec" "$out" <<'PY'
import json,sys
d=json.load(open(sys.argv[2])); k=d["roofline"]["kernels"]
print("%-40s ms/step %7.3f frac %.4f | "%(sys.argv[1], d["ms_per_step"], d["roofline"]["frac"]) + " ".join("%s=%.3f"%(n.replace("slice_fft","sf").replace("overlap_add","ola"),v["ms_per_step"]) for n,v in k.items()), "err %.1e"%d["max_abs_err_target0"])
PY
done
