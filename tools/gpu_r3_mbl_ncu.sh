#!/bin/bash
# ncu --set full of the MultiBoxLoss forward kernels (one launch each, after warm-up); the report is read on the build machine
mkdir -p gpurun_out
python bench_extra.py multibox > gpurun_out/mbl_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:'k_match_loss|k_mine_apply2|k_mbl_prepare' -s 30 -c 3 -f -o gpurun_out/r3_mbl_full python bench_extra.py multibox > gpurun_out/mbl_ncu_full.log 2>&1
echo "ncu rc $?"
