#!/bin/bash
# Capture the five sliCQT kernels once each with ncu --set full and leave CSV summaries in gpurun_out/.
# usage: tools/ncu_capture.sh <tag>
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file $OUT/launches_${TAG}.csv python tools/prof_step.py --batch 1 --steps 2 > $OUT/ncu_l.log 2>&1
echo "launch list rc=$?"
ncu --profile-from-start off --set full --cache-control none --clock-control none --import-source on -k regex:'slice_fft' -c 3 \
    -o $OUT/prof_${TAG}_slice python tools/prof_step.py --batch 2 --steps 1 > $OUT/ncu_f1.log 2>&1
echo "slice capture rc=$?"
ncu --profile-from-start off --set full --cache-control none --clock-control none -k regex:'bins_' -c 2 \
    -o $OUT/prof_${TAG}_bins python tools/prof_step.py --batch 2 --steps 1 > $OUT/ncu_f2.log 2>&1
echo "bins capture rc=$?"
for f in slice bins; do
  ncu -i $OUT/prof_${TAG}_$f.ncu-rep --page raw --csv > $OUT/prof_${TAG}_$f.raw.csv 2>/dev/null
done
ncu -i $OUT/prof_${TAG}_slice.ncu-rep --page source --csv > $OUT/prof_${TAG}_slice.source.csv 2>/dev/null
ls -la $OUT
du -sm $OUT
