#!/bin/bash
# (run under gpurun) ncu --set full of the four kernels with source correlation, aggregated per CUDA source line on the box
# (the full cuda,sass source page of the bin kernels is > 64 MB: only the aggregate travels back)
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:'slice_fft|bins_' -c 6 -o /tmp/prof_lines python tools/prof_step.py --batch 8 --steps 1 > gpurun_out/ncu_lines.log 2>&1
ncu -i /tmp/prof_lines.ncu-rep --page source --csv --print-source cuda,sass 2>/dev/null | python tools/ncu_lines.py "" 40 > gpurun_out/r2_source_lines.txt
head -3 gpurun_out/r2_source_lines.txt | cut -c1-200
