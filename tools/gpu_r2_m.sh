#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multibox_gpu.py tests/test_detect_paths_gpu.py -m gpu -x -q > gpurun_out/r2m_pytest.log 2>&1
echo "pytest rc $?"; tail -3 gpurun_out/r2m_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2m_bench.log 2> gpurun_out/r2m_bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2m_bench.log").read().strip().splitlines()[-1])
print("value %.0f ms/step %.5f lat %.4f frac %.3f e2e %.0f cpu %s" % (d["value"], d["ms_per_step"], d["latency"]["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"], d.get("cpu_baseline",{}).get("value")))
for k,v in (d.get("secondary") or {}).items(): print("   ", k, v.get("ms"), v.get("roofline_frac"), v.get("parity"), v.get("error"))
PY
python tools/multibox_once.py 3 > gpurun_out/r2m_mb_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2m_mb_launches.csv python tools/multibox_once.py 3 > gpurun_out/r2m_mb_list.log 2>&1
