#!/bin/bash
N=2
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 tests/multigpu_check.py > gpurun_out/r2k_check_$N.log 2>&1
echo "multigpu_check rc $?"; grep -E "CHECK FAILED|MULTIGPU_CHECK|sharded " gpurun_out/r2k_check_$N.log | head -20
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $N --steps 100 --warmup 5 > gpurun_out/r2k_bench_2.log 2> gpurun_out/r2k_bench_2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus $N --steps 100 --warmup 5 --mode clustered > gpurun_out/r2k_bench_2_clustered.log 2> gpurun_out/r2k_bench_2_clustered.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29614 tools/peer_breakdown.py 100 > gpurun_out/r2k_peer_breakdown_2.txt 2>&1
grep -v -i "warn\|OMP\|\*\*\*" gpurun_out/r2k_peer_breakdown_2.txt | tail -7
timeout 900 python -m pytest tests/test_detect_paths_gpu.py tests/test_detect_gpu.py -m gpu -x -q > gpurun_out/r2k_pytest.log 2>&1; tail -2 gpurun_out/r2k_pytest.log
timeout 300 python bench.py --steps 100 --warmup 5 --no-secondary --no-cpu-baseline > gpurun_out/r2k_bench_1.log 2> gpurun_out/r2k_bench_1.err
for f in gpurun_out/r2k_bench_*.log; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "n %d value %.0f ms/step %.4f lat %.4f gather_check %s" % (d["n_gpus"], d["value"], d["ms_per_step"], d["latency"]["ms_per_step"], d.get("gather_check")))
except Exception as e:
    print("  parse error", e)
PY
done
