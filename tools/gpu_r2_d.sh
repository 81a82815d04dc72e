#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_detect_paths_gpu.py tests/test_detect_gpu.py tests/test_heads.py tests/test_siblings.py tests/test_property_gpu.py -m gpu -x -q > gpurun_out/r2h_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2h_pytest.log
tail -3 gpurun_out/r2h_pytest.log
timeout 300 python bench.py --steps 100 --warmup 5 --depth 4 --no-secondary --no-cpu-baseline > gpurun_out/r2h_bench_d4.log 2> gpurun_out/r2h_bench_d4.err
timeout 300 python tools/k3_steady_profile.py 60 3 > gpurun_out/r2h_steady.txt 2>&1
cat gpurun_out/r2h_steady.txt
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2h_bench_full.log 2> gpurun_out/r2h_bench_full.err
FDT_DETECT_FUSED=0 timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2h_bench_full_unfused.log 2> gpurun_out/r2h_bench_full_unfused.err
for f in gpurun_out/r2h_bench_*.log; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=d.get("roofline") or {}
    print("  value %.0f ms/step %.4f lat %.4f frac %.3f e2e %.0f" % (d["value"], d["ms_per_step"], d["latency"]["ms_per_step"], r.get("frac",0), d["e2e"]["value"]))
    for k,v in (d.get("secondary") or {}).items(): print("   ", k, v.get("ms"), v.get("roofline_frac"), v.get("parity"), v.get("error"))
except Exception as e:
    print("  parse error", e)
PY
done
