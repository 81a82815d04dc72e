#!/bin/bash
# quick check of a Detect kernel change: the Detect GPU tests, then bench at the driver's K and at K = 100 (three runs each)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_detect_paths_gpu.py tests/test_detect_gpu.py -m gpu -x -q > gpurun_out/q_pytest.log 2>&1
echo "pytest rc $?"; tail -2 gpurun_out/q_pytest.log
for i in 1 2 3; do
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-secondary --no-cpu-baseline > gpurun_out/q_k20_$i.log 2> gpurun_out/q_k20_$i.err
timeout 600 python bench.py --steps 100 --warmup 5 --no-secondary --no-cpu-baseline > gpurun_out/q_k100_$i.log 2> gpurun_out/q_k100_$i.err
done
for f in gpurun_out/q_k20_*.log gpurun_out/q_k100_*.log; do python - "$f" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], "value %.0f ms/step %.5f" % (d["value"], d["ms_per_step"]), "frac", (d.get("roofline") or {}).get("frac"), "e2e", d["e2e"]["value"])
PY
done
