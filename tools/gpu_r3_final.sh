#!/bin/bash
# verification of the tree: smoke, all GPU tests, the driver's bench commands, MultiBoxLoss launch list + ncu summary
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r3_smoke.log 2>&1; echo "smoke rc $?"; tail -1 gpurun_out/r3_smoke.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r3_pytest.log 2>&1
echo "pytest rc $?"; tail -3 gpurun_out/r3_pytest.log
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r3_bench_k20.log 2> gpurun_out/r3_bench_k20.err
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r3_bench_ref.log 2> gpurun_out/r3_bench_ref.err
timeout 600 python bench.py --steps 100 --warmup 5 --no-secondary --no-cpu-baseline > gpurun_out/r3_bench_k100.log 2> gpurun_out/r3_bench_k100.err
for f in gpurun_out/r3_bench_k20.log gpurun_out/r3_bench_k100.log gpurun_out/r3_bench_ref.log; do python - "$f" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], "value %.0f ms/step %.5f" % (d["value"], d["ms_per_step"]), "frac", (d.get("roofline") or {}).get("frac"), "e2e", d["e2e"]["value"])
for k,v in (d.get("secondary") or {}).items(): print("   ", k, v.get("ms"), v.get("roofline_frac"), v.get("parity"), v.get("error"))
PY
done
python bench_extra.py multibox > gpurun_out/r3_mbl_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:'k_match|k_loss|k_mine|k_mbl|k_multibox' -c 24 --csv --log-file gpurun_out/r3_mbl_launches.csv python bench_extra.py multibox > gpurun_out/r3_mbl_ncu.log 2>&1
echo "mbl list rc $?"
ncu --set full --clock-control none --import-source on -k regex:'k_match_loss|k_mine_apply2|k_mbl_prepare' -s 30 -c 3 -f -o gpurun_out/r3_mbl_full python bench_extra.py multibox > gpurun_out/r3_mbl_ncu_full.log 2>&1
echo "mbl full rc $?"
