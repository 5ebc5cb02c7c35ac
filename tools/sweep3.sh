#!/bin/bash
# usage (under gpurun): tools/sweep3.sh "<variant>[,ENV=VAL...]" ...
for spec in "$@"; do
  IFS=',' read -ra parts <<< "$spec"
  v=${parts[0]}; envs=("${parts[@]:1}")
  lib=$PWD/build/variants/$v/libslicq.so
  [ "$v" = base ] && lib=$PWD/xumx_slicq_b200/libslicq.so
  out=gpurun_out/sw_${spec//[^A-Za-z0-9_=]/_}.json
  env SLICQ_B200_LIB=$lib "${envs[@]}" python bench.py --steps 8 --warmup 3 --no-cpu-baseline > $out 2>gpurun_out/sw_err.log || { echo "$spec FAILED"; tail -3 gpurun_out/sw_err.log; continue; }
  python - "$spec" "$out" <<'PY'
import json,sys
d=json.load(open(sys.argv[2])); k=d["roofline"]["kernels"]
print("%-36s ms/step %7.3f frac %.4f | "%(sys.argv[1], d["ms_per_step"], d["roofline"]["frac"]) + " ".join("%s=%.3f"%(n.replace("slice_fft","sf"),v["ms_per_step"]) for n,v in k.items()), "err %.1e"%d["max_abs_err_target0"])
PY
done
