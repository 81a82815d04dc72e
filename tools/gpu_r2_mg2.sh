#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 tests/multigpu_check.py > gpurun_out/r2mg2_check_$N.log 2>&1
echo "multigpu_check rc $?"; grep -E "CHECK FAILED|MULTIGPU_CHECK" gpurun_out/r2mg2_check_$N.log | head -20
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2mg2_pytest.log 2>&1
echo "pytest rc $?"; tail -3 gpurun_out/r2mg2_pytest.log
timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/r2mg2_bench_1.log 2> gpurun_out/r2mg2_bench_1.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2mg2_bench_1.log").read().strip().splitlines()[-1])
print("  value %.0f ms/step %.4f" % (d["value"], d["ms_per_step"]))
for k,v in (d.get("secondary") or {}).items(): print("   ", k, v.get("ms"), v.get("roofline_frac"), v.get("parity"), v.get("error"))
PY
