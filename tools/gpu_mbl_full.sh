#!/bin/bash
# ncu --set full of the MultiBoxLoss forward kernels (config 3), one launch each
mkdir -p gpurun_out
CMD="python bench_extra.py multibox"
$CMD > gpurun_out/mbl_plain.log 2>&1; echo "plain rc $?"
ncu --set full --clock-control none --import-source on -k regex:'k_match_default|k_loss_prior|k_mine_select|k_mine_apply' -s 8 -c 4 -o gpurun_out/mbl_full $CMD > gpurun_out/mbl_full.log 2>&1
echo "ncu rc $?"; ls -la gpurun_out/mbl_full.ncu-rep
