#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2i_pytest.log 2>&1
echo "pytest rc $?"; tail -4 gpurun_out/r2i_pytest.log
timeout 300 python bench.py --steps 100 --warmup 5 --no-secondary --no-cpu-baseline > gpurun_out/r2i_bench.log 2> gpurun_out/r2i_bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2i_bench.log").read().strip().splitlines()[-1])
print("  value %.0f ms/step %.4f lat %.4f frac %.3f" % (d["value"], d["ms_per_step"], d["latency"]["ms_per_step"], d["roofline"]["frac"]))
PY
timeout 300 python tools/k3_steady_profile.py 60 4 > gpurun_out/r2i_steady.txt 2>&1; cat gpurun_out/r2i_steady.txt
python tools/multibox_once.py 3 > gpurun_out/r2i_mb_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2i_mb_launches.csv python tools/multibox_once.py 3 > gpurun_out/r2i_mb_list.log 2>&1
python tools/multibox_once.py 3 > gpurun_out/r2i_mb_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_loss_prior|k_match_default|k_mine_select|k_mine_apply" -s 12 -c 4 -o gpurun_out/r2i_prof_multibox python tools/multibox_once.py 3 > gpurun_out/r2i_mb_full.log 2>&1
echo "ncu rc $?"
