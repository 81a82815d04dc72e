#!/bin/bash
# multi-GPU pass: N = number of GPUs of the box
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 tests/multigpu_check.py > gpurun_out/r2mg_check_$N.log 2>&1
echo "multigpu_check rc $?"; tail -3 gpurun_out/r2mg_check_$N.log
for g in peer peer-all nccl; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $N --steps 100 --warmup 5 --gather $g > gpurun_out/r2mg_bench_${N}_$g.log 2> gpurun_out/r2mg_bench_${N}_$g.err
echo "bench $g rc $?"
done
timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/r2mg_bench_1.log 2> gpurun_out/r2mg_bench_1.err
for f in gpurun_out/r2mg_bench_*.log; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("  value %.0f ms/step %.4f lat %.4f gather_check %s e2e %.0f" % (d["value"], d["ms_per_step"], d["latency"]["ms_per_step"], d.get("gather_check"), d["e2e"]["value"]))
    for k,v in (d.get("secondary") or {}).items(): print("   ", k, v.get("ms"), v.get("roofline_frac"), v.get("parity"), v.get("error"))
except Exception as e:
    print("  parse error", e)
PY
done
