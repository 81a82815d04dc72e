#!/bin/bash
# MultiBoxLoss change check: its GPU tests, then the forward timing (three runs) and the per-kernel launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multibox_gpu.py -m gpu -x -q > gpurun_out/m_pytest.log 2>&1
echo "pytest rc $?"; tail -3 gpurun_out/m_pytest.log
for i in 1 2 3; do python bench_extra.py multibox 2>&1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])['results']
print({k:(round(v['forward_ms']*1e3,1), round(v['module_fwd_bwd_ms']*1e3,1)) for k,v in d.items()})"; done
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:'k_match|k_loss|k_mine|k_conf|k_multibox' -c 12 --csv --log-file gpurun_out/mbl_launches.csv python bench_extra.py multibox > gpurun_out/mbl_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/mbl_launches.csv')) if len(r)>10]
ix={h:i for i,h in enumerate(rows[0])}
d=collections.OrderedDict()
for r in rows[1:]:
    d.setdefault((r[ix['ID']], r[ix['Kernel Name']][:36]),{})[r[ix['Metric Name']][:12]]=float(r[ix['Metric Value']].replace(',',''))
for i,(k,v) in enumerate(d.items()):
    if i<6: print(k, {a:round(b/(1e3 if 'time' in a else 1e6),2) for a,b in v.items()})
PY
