#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --steps 10 --warmup 3 --no-secondary --no-cpu-baseline"
$CMD > gpurun_out/r2_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv $CMD > gpurun_out/r2_ncu_list.log 2>&1
echo "launch list rc $?"
$CMD > gpurun_out/r2_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_sort_nms -s 20 -c 2 -o gpurun_out/r2_prof_sortnms_fused $CMD > gpurun_out/r2_ncu_full.log 2>&1
echo "full rc $?"
ls -la gpurun_out/r2_*
