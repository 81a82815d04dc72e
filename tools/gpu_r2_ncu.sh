#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --steps 10 --warmup 3 --no-secondary --no-cpu-baseline"
$CMD > gpurun_out/r2f_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/r2f_launches.csv $CMD > gpurun_out/r2f_ncu_list.log 2>&1
echo "launch list rc $?"
$CMD > gpurun_out/r2f_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_sort_nms -s 30 -c 3 -o gpurun_out/r2f_prof_sortnms_fused $CMD > gpurun_out/r2f_ncu_full.log 2>&1
echo "full rc $?"
tail -3 gpurun_out/r2f_ncu_full.log
ls -la gpurun_out/r2f_*
