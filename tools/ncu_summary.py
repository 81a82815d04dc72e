"""Turns one ncu launch list (--metrics gpu__time_duration.sum --csv) and one `ncu --set full` report of the dominant kernel into the
small summary bench.py reads (profiles/rNN_detect_summary.json): DRAM bytes per launch of the dominant kernel and its share of the
step's kernel time.

    python tools/ncu_summary.py <launches.csv> <prof.ncu-rep> <kernel-substring> <out.json> <order> [note]
"""
import collections
import csv
import json
import subprocess
import sys

launches, rep, kern, out_path, order = sys.argv[1:6]
note = sys.argv[6] if len(sys.argv) > 6 else ""

# ---- launch list: device time per kernel name (cold-cache, serialised: shares only)
rows = [r for r in csv.reader(open(launches, errors="replace")) if len(r) > 5]
hdr, per = None, collections.defaultdict(list)
for r in rows:
    if r[0] == "ID":
        hdr = r
        continue
    if hdr is None:
        continue
    d = dict(zip(hdr, r))
    try:
        v = float(d["Metric Value"])
    except ValueError:
        continue
    unit = d["Metric Unit"]
    per[d["Kernel Name"]].append(v * {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3}.get(unit, 1e-3))
ours = {k: v for k, v in per.items() if "k_" in k and ("<unnamed>::k_" in k or "fdt" in k)}
dom = [k for k in ours if kern in k and "(bool)1" in k or kern in k and ", 1>" in k]
dom = dom or [k for k in ours if kern in k]
step_kernels = {k: v for k, v in ours.items() if k in dom or "k_detect_begin" in k}
n_dom = sum(len(per[k]) for k in dom)
t_dom = sum(sum(per[k]) for k in dom)
# the step of the timed path = one k_detect_begin + one dominant kernel: compare means
mean_dom = t_dom / max(n_dom, 1)
begin = [k for k in ours if "k_detect_begin" in k]
mean_begin = sum(sum(per[k]) for k in begin) / max(sum(len(per[k]) for k in begin), 1)
share = mean_dom / (mean_dom + mean_begin)

# ---- full report: DRAM traffic and a few headline counters of the dominant kernel
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
h, units = rr[0], rr[1]


def col(name):
    return h.index(name)


def to_bytes(v, u):
    return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]


caps = [r for r in rr[2:] if kern in r[col("Kernel Name")]]
rd = sum(to_bytes(r[col("dram__bytes_read.sum")], units[col("dram__bytes_read.sum")]) for r in caps) / len(caps)
wr = sum(to_bytes(r[col("dram__bytes_write.sum")], units[col("dram__bytes_write.sum")]) for r in caps) / len(caps)
inst = sum(float(r[col("smsp__inst_executed.sum")]) for r in caps) / len(caps)
issue = sum(float(r[col("smsp__issue_active.avg.pct_of_peak_sustained_active")]) for r in caps) / len(caps)
dur = sum(float(r[col("gpu__time_duration.sum")]) for r in caps) / len(caps)
summary = {"order": int(order), "kernel": caps[0][col("Kernel Name")][:90], "captures": len(caps),
           "dram_bytes_read": rd, "dram_bytes_write": wr,
           "traffic_note": f"ncu --set full, {len(caps)} launches: dram__bytes_read.sum {rd / 1e6:.2f} MB + dram__bytes_write.sum {wr / 1e3:.1f} KB per launch "
                           f"(output rows are written back from L2 after the kernel)",
           "warp_instructions": inst, "issue_active_pct": issue, "duration_us_under_ncu": dur,
           "registers": int(float(caps[0][col("launch__registers_per_thread")])),
           "kernel_share_of_step": share,
           "share_note": f"ncu launch list (cold-cache, serialised): dominant kernel {mean_dom:.1f} us x {n_dom}, k_detect_begin {mean_begin:.1f} us per launch",
           "launch_list_us": {k[:70]: {"n": len(v), "mean_us": sum(v) / len(v)} for k, v in sorted(ours.items(), key=lambda kv: -sum(kv[1]))},
           "note": note}
json.dump(summary, open(out_path, "w"), indent=1)
print(json.dumps(summary, indent=1))
