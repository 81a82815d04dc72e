#!/bin/bash
# round 2, first GPU pass: all GPU tests, then the headline bench at several workspace depths / kernel variants
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2a_pytest.log
for d in 1 2 3 4; do
  timeout 300 python bench.py --steps 100 --warmup 5 --depth $d --no-secondary --no-cpu-baseline > gpurun_out/r2a_bench_d$d.log 2> gpurun_out/r2a_bench_d$d.err
done
FDT_K3_CLUSTER=1 timeout 300 python bench.py --steps 100 --warmup 5 --depth 3 --no-secondary --no-cpu-baseline > gpurun_out/r2a_bench_d3_cl2.log 2> gpurun_out/r2a_bench_d3_cl2.err
FDT_K3_CLUSTER=0 timeout 300 python bench.py --steps 100 --warmup 5 --depth 1 --no-secondary --no-cpu-baseline > gpurun_out/r2a_bench_d1_cl1.log 2> gpurun_out/r2a_bench_d1_cl1.err
timeout 300 python bench.py --steps 20 --warmup 3 --depth 3 --no-secondary --no-cpu-baseline > gpurun_out/r2a_bench_d3_k20.log 2> gpurun_out/r2a_bench_d3_k20.err
timeout 600 python bench.py > gpurun_out/r2a_bench_full.log 2> gpurun_out/r2a_bench_full.err
tail -3 gpurun_out/r2a_pytest.log
for f in gpurun_out/r2a_bench_*.log; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=d.get("roofline") or {}
    print("  value %.0f ms/step %.4f lat %.4f k3 b2b %s iso %s e2e %.0f" % (d["value"], d["ms_per_step"], d["latency"]["ms_per_step"], r.get("kernel_ms"), r.get("kernel_ms_isolated"), d["e2e"]["value"]))
except Exception as e:
    print("  parse error", e)
PY
done
