#!/bin/bash
# (run under gpurun) ncu --set full of the bin kernels with source correlation, aggregated per CUDA source line on the box
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:bins_ -c 4 -o /tmp/prof_bins python tools/prof_step.py --batch 8 --steps 1 > gpurun_out/ncu_bins.log 2>&1
ncu -i /tmp/prof_bins.ncu-rep --page source --csv --print-source cuda,sass 2>/dev/null | python tools/ncu_lines.py bins_ 70 > gpurun_out/bins_lines.txt
ncu -i /tmp/prof_bins.ncu-rep --page raw --csv > gpurun_out/prof_bins.raw.csv
head -5 gpurun_out/bins_lines.txt | cut -c1-200
