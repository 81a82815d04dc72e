"""Joins an `ncu --page source --csv` SASS export with nvdisasm line info: instructions / stall samples per source line.
usage: python tools/sass_lines.py <src.csv> <nvdisasm -g -c output> <mangled-kernel-substring> [top]"""
import csv
import re
import sys
from collections import defaultdict

src_csv, sass, kern = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 60
# nvdisasm: per function, instruction offset -> (file line, inlined-at chain top line)
lines = {}
cur_fn, cur_line, in_fn = None, None, False
for ln in open(sass, errors="replace"):
    m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
    if m:
        in_fn = kern in m.group(1)
        cur_line = None
        continue
    if not in_fn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', ln)
    if m:
        # with `nvdisasm -gi` an inline chain is printed innermost first; keep the outermost frame (the kernel body)
        if m.group(3) and not __import__("os").environ.get("INNER"):
            cur_line = (m.group(3).split("/")[-1], int(m.group(4)), "")
        else:
            cur_line = (m.group(1).split("/")[-1], int(m.group(2)), "")
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m:
        lines[int(m.group(1), 16)] = (cur_line, m.group(2).strip())
rows = list(csv.reader(open(src_csv)))
h = rows[1]
ci, cs, ca = h.index("Instructions Executed"), h.index("# Samples"), h.index("Address")
cth = h.index("Thread Instructions Executed")
base = None
per = defaultdict(lambda: [0, 0, 0])
tot_i = tot_s = 0
for r in rows[2:]:
    if len(r) <= ci or not r[ci].isdigit():
        continue
    a = int(r[ca], 16)
    if base is None:
        base = a
    key = lines.get(a - base, (None, ""))[0]
    k = (key[0], key[1]) if key else ("?", 0)
    per[k][0] += int(r[ci]); per[k][1] += int(r[cs]); per[k][2] += int(r[cth])
    tot_i += int(r[ci]); tot_s += int(r[cs])
print(f"total warp instructions {tot_i}, samples {tot_s}")
import os
if os.environ.get("BINS"):
    # BINS="name:lo-hi,name:lo-hi": instructions / samples per source-line range of detect.cu
    for spec in os.environ["BINS"].split(","):
        name, rng = spec.split(":"); lo, hi = map(int, rng.split("-"))
        i = sum(v[0] for (f, l), v in per.items() if f == "detect.cu" and lo <= l <= hi)
        sm = sum(v[1] for (f, l), v in per.items() if f == "detect.cu" and lo <= l <= hi)
        t = sum(v[2] for (f, l), v in per.items() if f == "detect.cu" and lo <= l <= hi)
        print(f"  {name:18s} lines {lo}-{hi}: instr {i:>9} {100 * i / tot_i:5.1f}%  samples {100 * sm / max(tot_s, 1):5.1f}%  lanes {t / max(i, 1):4.1f}")

srcs = {}
for (f, l), (i, s, t) in sorted(per.items(), key=lambda kv: -kv[1][0])[:top]:
    if f not in srcs:
        try:
            srcs[f] = open(f"/root/repo/face-detection-and-tracking_b200/csrc/{f}").read().split("\n")
        except OSError:
            srcs[f] = []
    text = srcs[f][l - 1].strip()[:100] if 0 < l <= len(srcs[f]) else ""
    print(f"{f}:{l:<5} instr {i:>9} {100 * i / tot_i:5.1f}%  samples {100 * s / max(tot_s, 1):5.1f}%  lanes {t / max(i, 1):4.1f}  {text}")
