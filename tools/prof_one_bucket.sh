#!/bin/bash
# ncu source-level capture of the bins kernels restricted to one bucket. usage: tools/prof_one_bucket.sh <bucket> <tag>
B=$1; TAG=$2
export SLICQ_ONLY_BUCKET=$B SLICQ_BINS_JOBS=200000
ncu --profile-from-start off --set full --cache-control none --clock-control none --import-source on -k regex:'bins_' -c 2 \
    -o gpurun_out/prof_${TAG} python tools/prof_step.py --batch 16 --steps 1 > gpurun_out/ncu_one.log 2>&1
echo "rc=$?"
ncu -i gpurun_out/prof_${TAG}.ncu-rep --page raw --csv > gpurun_out/prof_${TAG}.raw.csv 2>/dev/null
ncu -i gpurun_out/prof_${TAG}.ncu-rep --page source --csv > gpurun_out/prof_${TAG}.source.csv 2>/dev/null
rm -f gpurun_out/prof_${TAG}.ncu-rep
ls -la gpurun_out | tail -5
