#!/usr/bin/env python
"""stdin: `ncu --page source --csv --print-source sass` of one kernel.  stdout: executed-instruction mix by opcode, share of
instructions / stall samples by active-lane count, stall reasons, and the 25 most-stalled instructions."""
import csv, sys
from collections import Counter
rows = list(csv.reader(sys.stdin))
if len(rows) < 3:
    sys.exit("no source page")
print(rows[0][1] if len(rows[0]) > 1 else rows[0])
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
I = lambda r, k: int(r[ix[k]] or 0)
tot_s = sum(I(r, '# Samples') for r in data) or 1
tot_i = sum(I(r, 'Instructions Executed') for r in data) or 1
print(f"{len(data)} SASS instructions; {tot_i} warp-instructions executed; {tot_s} stall samples")
cs, ci = Counter(), Counter()
for r in data:
    parts = r[ix['Source']].split()
    op = parts[1] if parts[0].startswith('@') else parts[0]
    op = '.'.join(op.split('.')[:2]) if op.startswith(('IMAD', 'I2F', 'F2I', 'MUFU')) else op.split('.')[0]
    cs[op] += I(r, '# Samples'); ci[op] += I(r, 'Instructions Executed')
print("\nopcode            executed      share   stall-sample share")
for op, n in ci.most_common(28):
    print(f"{op:14s} {n:12d}  {100 * n / tot_i:6.1f}%  {100 * cs[op] / tot_s:6.1f}%")
bins = {}
for r in data:
    n = I(r, 'Instructions Executed')
    if not n:
        continue
    b = min(int(float(r[ix['Avg. Threads Executed']]) // 4) * 4, 32)
    a = bins.setdefault(b, [0, 0]); a[0] += n; a[1] += I(r, '# Samples')
print("\nactive lanes   instr share   sample share")
for b in sorted(bins):
    print(f"  {b:2d}-{min(b + 3, 32):2d}       {100 * bins[b][0] / tot_i:6.1f}%      {100 * bins[b][1] / tot_s:6.1f}%")
print("\nstall reasons (share of samples)")
for k in hdr:
    if k.startswith('stall_') and 'Not' not in k:
        s = sum(I(r, k) for r in data)
        if s > tot_s * 0.01:
            print(f"  {k[6:]:20s} {100 * s / tot_s:5.1f}%")
print("\nmost-stalled instructions: samples, executions, avg lanes, SASS, top stall reasons")
for i in sorted(range(len(data)), key=lambda i: -I(data[i], '# Samples'))[:25]:
    r = data[i]
    st = {k[6:]: I(r, k) for k in hdr if k.startswith('stall_') and 'Not' not in k and I(r, k) > 0}
    top = sorted(st.items(), key=lambda kv: -kv[1])[:3]
    print(f"  {I(r, '# Samples'):6d} {I(r, 'Instructions Executed'):9d} {r[ix['Avg. Threads Executed']]:>5s}  {r[ix['Source']].strip()[:64]:64s} {top}")
