#!/bin/bash
# 8-GPU box: multigpu_check at 8 ranks, then the scaling points N = 4, 8 (weak scaling, 64 images per GPU) and the per-variant breakdown
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 tests/multigpu_check.py > gpurun_out/r2s_check_8.log 2>&1
echo "multigpu_check(8) rc $?"; grep -E "CHECK FAILED|MULTIGPU_CHECK|sharded " gpurun_out/r2s_check_8.log | head -20
for N in 8 4; do
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2961$N bench.py --gpus $N --steps 100 --warmup 5 > gpurun_out/r2s_bench_$N.log 2> gpurun_out/r2s_bench_$N.err
  echo "bench $N rc $?"
done
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29620 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r2s_bench_8_k20.log 2> gpurun_out/r2s_bench_8_k20.err
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29622 tools/peer_breakdown.py 100 > gpurun_out/r2s_peer_breakdown_8.txt 2>&1
cat gpurun_out/r2s_peer_breakdown_8.txt | grep -v Warning | tail -8
for f in gpurun_out/r2s_bench_*.log; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("  n %d value %.0f ms/step %.4f lat %.4f gather_check %s e2e %.0f blocks %s" % (d["n_gpus"], d["value"], d["ms_per_step"], d["latency"]["ms_per_step"], d.get("gather_check"), d["e2e"]["value"], [round(x,3) for x in d["timing"]["block_ms"]]))
except Exception as e:
    print("  parse error", e)
PY
done
