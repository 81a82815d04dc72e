#!/bin/bash
# ncu --set full of the MultiBoxLoss forward kernels (one launch each, after warm-up), summarised to text
mkdir -p gpurun_out
python bench_extra.py multibox > gpurun_out/mbl_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:'k_match_loss|k_mine_apply2|k_mbl_prepare' -s 30 -c 3 -f -o gpurun_out/r3_mbl_full python bench_extra.py multibox > gpurun_out/mbl_ncu_full.log 2>&1
echo "ncu rc $?"
ncu -i gpurun_out/r3_mbl_full.ncu-rep --page raw --csv > gpurun_out/r3_mbl_full_raw.csv 2>/dev/null
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/r3_mbl_full_raw.csv')))
hdr=rows[0]
want=['Kernel Name','gpu__time_duration.sum','smsp__inst_executed.sum','sm__inst_executed_pipe_fp64.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__warps_active.avg.pct_of_peak_sustained_active','achieved_occupancy','smsp__average_warp_latency_issue_stalled_barrier','dram__bytes_read.sum','dram__bytes_write.sum','launch__registers_per_thread','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','launch__occupancy_limit_warps','sm__cycles_active.avg','sm__cycles_elapsed.max']
stall=[h for h in hdr if 'smsp__average_warps_issue_stalled' in h and 'not_issued' not in h and h.endswith('_per_issue_active.ratio')] or [h for h in hdr if 'issue_stalled' in h and h.endswith('ratio')]
for r in rows[2:]:
    d=dict(zip(hdr,r))
    print('=====', d.get('Kernel Name','')[:50])
    for w in want:
        if w in d: print('  ', w, d[w])
    ss=sorted(((float(d[h].replace(',','')) if d[h] not in ('','n/a') else 0.0, h) for h in stall), reverse=True)[:8]
    for v,h in ss: print('   stall', h.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio',''), round(v,2))
PY
