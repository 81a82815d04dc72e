#!/bin/bash
mkdir -p gpurun_out
python tools/mbl_diff.py 16 98 0 40 > gpurun_out/r2j_mbl_diff.txt 2>&1; tail -12 gpurun_out/r2j_mbl_diff.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2j_pytest.log 2>&1
echo "pytest rc $?"; tail -4 gpurun_out/r2j_pytest.log
timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/r2j_bench.log 2> gpurun_out/r2j_bench.err
timeout 300 python bench.py --steps 100 --warmup 5 --mode clustered --no-secondary --no-cpu-baseline > gpurun_out/r2j_bench_clustered.log 2> gpurun_out/r2j_bench_clustered.err
for f in gpurun_out/r2j_bench*.log; do python - "$f" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], "  value %.0f ms/step %.4f lat %.4f frac %.3f" % (d["value"], d["ms_per_step"], d["latency"]["ms_per_step"], d["roofline"]["frac"]))
for k,v in (d.get("secondary") or {}).items(): print("   ", k, v.get("ms"), v.get("roofline_frac"), v.get("parity"), v.get("error"))
PY
done
timeout 300 python tools/k3_steady_profile.py 60 4 > gpurun_out/r2j_steady.txt 2>&1; cat gpurun_out/r2j_steady.txt
