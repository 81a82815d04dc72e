#!/bin/bash
# per-kernel times of the MultiBoxLoss forward (config 3): plain run first, then the ncu launch list of the same command
mkdir -p gpurun_out
CMD="python bench_extra.py multibox"
$CMD > gpurun_out/mbl_plain.log 2>&1; echo "plain rc $?"; tail -5 gpurun_out/mbl_plain.log
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'k_match|k_loss|k_mine|k_conf|k_multibox' -c 120 --csv --log-file gpurun_out/mbl_launches.csv $CMD > gpurun_out/mbl_ncu.log 2>&1
echo "ncu rc $?"
