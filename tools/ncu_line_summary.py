#!/usr/bin/env python
"""stdin: `ncu --page source --csv --print-source cuda,sass` of one kernel compiled with -lineinfo.  stdout: executed
warp-instructions and stall samples per CUDA source line (the 50 heaviest): which lines of engine.cu / physics.cuh / rng.cuh the
kernel spends its issue slots on.  (The csv has one section per source file; a row with a line number carries the totals of the
SASS rows listed under it.)"""
import csv, sys
from collections import Counter
by_i, by_s, lanes = Counter(), Counter(), {}
fname, hdr = "?", None
for r in csv.reader(sys.stdin):
    if len(r) >= 2 and r[0] == "File Path":
        fname = r[1].split("/")[-1]; hdr = None; continue
    if len(r) >= 2 and r[0] == "Line No":
        hdr = r; continue
    if hdr is None or len(r) < 12 or not r[0].strip().isdigit():
        continue
    ie, ns, at = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Thread Instructions Executed")
    try:
        n, s, t = int(r[ie]), int(r[ns]), int(r[at])
    except ValueError:
        continue
    key = f"{fname}:{r[0]}  {r[1].strip()[:100]}"
    by_i[key] += n; by_s[key] += s; lanes[key] = lanes.get(key, 0) + t
tot_i, tot_s = sum(by_i.values()) or 1, sum(by_s.values()) or 1
print(f"{tot_i} warp-instructions, {tot_s} stall samples (attributed to source lines)")
for k, n in by_i.most_common(50):
    print(f"{100 * n / tot_i:5.1f}% instr {100 * by_s[k] / tot_s:5.1f}% samples  {lanes[k] / max(n, 1):4.1f} lanes  {k}")
