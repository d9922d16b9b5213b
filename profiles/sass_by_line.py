#!/usr/bin/env python
"""Join an ncu SASS source page (ncu -i X.ncu-rep --page source --csv) with nvdisasm line info
of the cubin, and print per-CUDA-source-line totals: warp instructions executed and stall samples.
usage: sass_by_line.py <source_page.csv> <cubin> <kernel-name-substring> [top N]"""
import csv, re, subprocess, sys
from collections import defaultdict

page, cubin, kname = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
SORT = 0 if (len(sys.argv) > 5 and sys.argv[5] == "inst") else 1
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
# split per function
line_of = {}
cur_fn, cur_line, on = None, None, False
for l in dis.splitlines():
    m = re.match(r"\s*\.section\s+\.text\.(\S+?),", l)
    if m:
        cur_fn = m.group(1); on = kname in cur_fn; continue
    if not on: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur_line = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        line_of[int(m.group(1), 16)] = cur_line
rows = list(csv.reader(open(page)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ia, ii, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
base = None
agg = defaultdict(lambda: [0, 0, defaultdict(int)])
tot_i = tot_s = 0
for r in rows[hi + 1:]:
    if len(r) < len(hdr): continue
    a = int(r[ia], 16)
    if base is None: base = a
    key = line_of.get(a - base, ("?", 0))
    n, s = int(r[ii] or 0), int(r[isamp] or 0)
    agg[key][0] += n; agg[key][1] += s
    for c in stall_cols:
        v = int(r[c] or 0)
        if v: agg[key][2][hdr[c]] += v
    tot_i += n; tot_s += s
print(f"total warp-instructions {tot_i}, samples {tot_s}")
for key, (n, s, st) in sorted(agg.items(), key=lambda kv: -kv[1][SORT])[:top]:
    tops = ", ".join(f"{k[6:]}={v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print(f"{key[0]}:{key[1]:<5d} inst {n:>10d} ({100*n/tot_i:5.1f}%)  samples {s:>7d} ({100*s/max(tot_s,1):5.1f}%)  {tops}")
