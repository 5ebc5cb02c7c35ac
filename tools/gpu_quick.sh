#!/bin/bash
# usage: tools/gpu_quick.sh <tag> [ncu-kernel-regex]   (run under gpurun): bench line + optional ncu capture of one kernel family
TAG=$1; PAT=$2
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err || tail -5 gpurun_out/${TAG}_bench.err
python - <<PY
import json
d=json.load(open('gpurun_out/${TAG}_bench.json'))
print("ms/step %.4f frac %.4f err %.2e"%(d['ms_per_step'], d['roofline']['frac'], d['max_abs_err_target0']), {k:v['ms_per_step'] for k,v in d['roofline']['kernels'].items()})
PY
if [ -n "$PAT" ]; then
  ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:$PAT -c 3 -o gpurun_out/prof_${TAG} python tools/prof_step.py --batch 4 --steps 1 > gpurun_out/ncu_${TAG}.log 2>&1
  ncu -i gpurun_out/prof_${TAG}.ncu-rep --page raw --csv > gpurun_out/prof_${TAG}.raw.csv
  ncu -i gpurun_out/prof_${TAG}.ncu-rep --page source --csv > gpurun_out/prof_${TAG}.source.csv
  rm -f gpurun_out/prof_${TAG}.ncu-rep
fi
