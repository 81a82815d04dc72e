#!/bin/bash
# 8-GPU box: multigpu_check at 8 ranks, then bench.py --gpus 8 as the driver launches it (K = 20) and with long blocks (K = 100)
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 tests/multigpu_check.py > gpurun_out/r3s_check_8.log 2>&1
echo "multigpu_check(8) rc $?"; grep -E "CHECK FAILED|MULTIGPU_CHECK" gpurun_out/r3s_check_8.log | head -20
for K in 20 100; do
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((29630 + K / 10)) bench.py --gpus 8 --steps $K --warmup 5 --no-secondary --no-cpu-baseline > gpurun_out/r3s_bench_8_k$K.log 2> gpurun_out/r3s_bench_8_k$K.err
echo "bench 8 K=$K rc $?"
python - gpurun_out/r3s_bench_8_k$K.log <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("  n %d value %.0f ms/step %.4f lat %.4f gather_check %s e2e %.0f" % (d["n_gpus"], d["value"], d["ms_per_step"], d["latency"]["ms_per_step"], d.get("gather_check"), d["e2e"]["value"]))
PY
done
