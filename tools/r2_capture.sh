#!/bin/bash
# Round-2 evidence capture (run under gpurun, 1 GPU): bench lines, ncu launch list, ncu --set full of the four kernels.
O=gpurun_out
python bench.py --steps 20 --warmup 5 > $O/r2_bench_n1.json 2> $O/r2_bench_n1.err || tail -3 $O/r2_bench_n1.err
python bench.py --impl reference --steps 20 --warmup 5 > $O/r2_bench_reference.json 2>/dev/null
python bench.py --mode slices --steps 20 --warmup 3 > $O/r2_bench_slices_n1.json 2>/dev/null
python bench.py --mode sweep --tracks 64 > $O/r2_bench_sweep64_n1.json 2>/dev/null
ncu --profile-from-start off --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__grid_size --clock-control none --csv --log-file $O/r2_launches.csv python tools/prof_step.py --batch 8 --steps 1 > $O/ncu_l.log 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:'slice_fft|bins_' -c 6 -o $O/prof_r2_final python tools/prof_step.py --batch 8 --steps 1 > $O/ncu_f.log 2>&1
ncu -i $O/prof_r2_final.ncu-rep --page raw --csv > $O/prof_r2_final.raw.csv
python tools/ncu_summary.py $O/prof_r2_final.raw.csv > $O/r2_ncu_summary.txt
python tools/ncu_stalls.py $O/prof_r2_final.raw.csv >> $O/r2_ncu_summary.txt
rm -f $O/prof_r2_final.ncu-rep
python -c "import json; d=json.load(open('$O/r2_bench_n1.json')); print('N1', d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['cpu_baseline']['value'])"
