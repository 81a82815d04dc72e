#!/bin/bash
# host-buffer pipeline: its tests, the older host tests, then the bench (e2e block)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_detect_host_gpu.py tests/test_detect_gpu.py -m gpu -x -q > gpurun_out/h_pytest.log 2>&1
echo "pytest rc $?"; tail -15 gpurun_out/h_pytest.log
for i in 1 2; do
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-secondary --no-cpu-baseline > gpurun_out/h_k20_$i.log 2> gpurun_out/h_k20_$i.err
tail -3 gpurun_out/h_k20_$i.err
python - gpurun_out/h_k20_$i.log <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
e=d["e2e"]; print("value %.0f ms/step %.5f" % (d["value"], d["ms_per_step"]), "e2e %.0f (%.3f ms) sync %.0f (%.3f ms) floor %.3f ms h2d %.1f GB/s" % (e["value"], e["ms_per_step"], e["sync_call"]["value"], e["sync_call"]["ms_per_step"], e["pcie_floor_ms"], e["pinned_h2d_gbs_this_host"]))
PY
done
for ch in 0 8 32; do
FDT_HOST_CHUNK=$ch timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-secondary --no-cpu-baseline > gpurun_out/h_chunk$ch.log 2> gpurun_out/h_chunk$ch.err
python - gpurun_out/h_chunk$ch.log $ch <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
e=d["e2e"]; print("chunk", sys.argv[2], "e2e %.0f (%.3f ms) sync %.0f (%.3f ms)" % (e["value"], e["ms_per_step"], e["sync_call"]["value"], e["sync_call"]["ms_per_step"]))
PY
done
