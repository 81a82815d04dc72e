#!/bin/bash
# round 2: GPU tests, then the headline bench on the fused path vs the two-kernel path, at several depths
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2b_pytest.log
for d in 1 2 3 4; do
  timeout 300 python bench.py --steps 100 --warmup 5 --depth $d --no-secondary --no-cpu-baseline > gpurun_out/r2b_bench_d$d.log 2> gpurun_out/r2b_bench_d$d.err
done
FDT_DETECT_FUSED=0 timeout 300 python bench.py --steps 100 --warmup 5 --depth 3 --no-secondary --no-cpu-baseline > gpurun_out/r2b_bench_d3_unfused.log 2> gpurun_out/r2b_bench_d3_unfused.err
timeout 300 python bench.py --steps 20 --warmup 3 --no-secondary --no-cpu-baseline > gpurun_out/r2b_bench_k20.log 2> gpurun_out/r2b_bench_k20.err
timeout 300 python bench.py --steps 100 --mode clustered --no-secondary --no-cpu-baseline > gpurun_out/r2b_bench_clustered.log 2> gpurun_out/r2b_bench_clustered.err
tail -3 gpurun_out/r2b_pytest.log
for f in gpurun_out/r2b_bench_*.log; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=d.get("roofline") or {}
    u=r.get("unfused_path") or {}
    print("  value %.0f ms/step %.4f lat %.4f frac %.3f unfused k3 b2b %s iso %s e2e %.0f" % (d["value"], d["ms_per_step"], d["latency"]["ms_per_step"], r.get("frac",0), u.get("k_sort_nms_ms_back_to_back"), u.get("k_sort_nms_ms_isolated"), d["e2e"]["value"]))
except Exception as e:
    print("  parse error", e)
PY
done
