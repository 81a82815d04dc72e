#!/bin/bash
# state check: smoke, all GPU tests, the driver's bench command, MultiBoxLoss per-kernel launch list
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r3_smoke.log 2>&1; echo "smoke rc $?"; tail -1 gpurun_out/r3_smoke.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r3_pytest.log 2>&1
echo "pytest rc $?"; tail -3 gpurun_out/r3_pytest.log
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r3_bench_k20.log 2> gpurun_out/r3_bench_k20.err
python - gpurun_out/r3_bench_k20.log <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], "value %.0f ms/step %.5f" % (d["value"], d["ms_per_step"]), "frac", (d.get("roofline") or {}).get("frac"), "e2e", d["e2e"]["value"])
for k,v in (d.get("secondary") or {}).items(): print("   ", k, v.get("ms"), v.get("roofline_frac"), v.get("parity"), v.get("error"))
PY
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:'k_match|k_loss|k_mine|k_conf|k_multibox' -c 12 --csv --log-file gpurun_out/mbl_launches.csv python bench_extra.py multibox > gpurun_out/mbl_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/mbl_launches.csv')) if len(r)>10]
ix={h:i for i,h in enumerate(rows[0])}
d=collections.OrderedDict()
for r in rows[1:]:
    d.setdefault((r[ix['ID']], r[ix['Kernel Name']][:36]),{})[r[ix['Metric Name']][:12]]=float(r[ix['Metric Value']].replace(',',''))
for i,(k,v) in enumerate(d.items()):
    if i<12: print(k, {a:round(b/(1e3 if 'time' in a else 1e6),2) for a,b in v.items()})
PY
