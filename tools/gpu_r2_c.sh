#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_detect_paths_gpu.py tests/test_detect_gpu.py tests/test_heads.py tests/test_siblings.py -m gpu -x -q > gpurun_out/r2c_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2c_pytest.log
tail -3 gpurun_out/r2c_pytest.log
for d in 3 4; do
  timeout 300 python bench.py --steps 100 --warmup 5 --depth $d --no-secondary --no-cpu-baseline > gpurun_out/r2c_bench_d$d.log 2> gpurun_out/r2c_bench_d$d.err
done
timeout 300 python tools/k3_steady_profile.py 60 3 > gpurun_out/r2c_steady.txt 2>&1
timeout 300 python tools/k3_steady_profile.py 60 1 > gpurun_out/r2c_steady_d1.txt 2>&1
cat gpurun_out/r2c_steady.txt gpurun_out/r2c_steady_d1.txt
for f in gpurun_out/r2c_bench_*.log; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=d.get("roofline") or {}
    print("  value %.0f ms/step %.4f lat %.4f frac %.3f e2e %.0f" % (d["value"], d["ms_per_step"], d["latency"]["ms_per_step"], r.get("frac",0), d["e2e"]["value"]))
except Exception as e:
    print("  parse error", e)
PY
done
