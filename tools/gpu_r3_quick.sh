#!/bin/bash
# quick check after a MultiBoxLoss / math change: math self-test, MultiBoxLoss tests, forward timing, per-kernel list, Detect bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_property_gpu.py tests/test_multibox_gpu.py -m gpu -x -q > gpurun_out/q_pytest.log 2>&1
echo "pytest rc $?"; tail -5 gpurun_out/q_pytest.log
for i in 1 2; do python bench_extra.py multibox 2>&1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])['results']
print({k:(round(v['forward_ms']*1e3,1), round(v['module_fwd_bwd_ms']*1e3,1)) for k,v in d.items()})"; done
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:'k_match|k_loss|k_mine|k_conf|k_multibox|k_mbl' -c 12 --csv --log-file gpurun_out/mbl_launches.csv python bench_extra.py multibox > gpurun_out/mbl_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/mbl_launches.csv')) if len(r)>10]
ix={h:i for i,h in enumerate(rows[0])}
d=collections.OrderedDict()
for r in rows[1:]:
    d.setdefault((r[ix['ID']], r[ix['Kernel Name']][:36]),{})[r[ix['Metric Name']][:12]]=float(r[ix['Metric Value']].replace(',',''))
for i,(k,v) in enumerate(d.items()):
    if i<7: print(k, {a:round(b/(1e3 if 'time' in a else 1e6),2) for a,b in v.items()})
PY
if [ "$1" == "detect" ]; then
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-secondary --no-cpu-baseline > gpurun_out/q_bench_k20.log 2> gpurun_out/q_bench_k20.err
timeout 600 python bench.py --gpus 1 --steps 100 --warmup 5 --no-secondary --no-cpu-baseline > gpurun_out/q_bench_k100.log 2> gpurun_out/q_bench_k100.err
for f in gpurun_out/q_bench_k20.log gpurun_out/q_bench_k100.log; do python - "$f" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], "value %.0f ms/step %.5f" % (d["value"], d["ms_per_step"]), "frac", (d.get("roofline") or {}).get("frac"), "lat", d["latency"]["ms_per_step"], "e2e", d["e2e"]["value"])
PY
done
fi
