#!/bin/bash
# SM-active share of every launch of one config-2 batch (all kernels): where the GPU idles because a few CTAs run long.
mkdir -p gpurun_out/balance
PB_SAMPLE_G=${G:-4} timeout -k 10 500 ncu --metrics gpu__time_duration.sum,sm__cycles_active.avg,sm__cycles_active.max,sm__cycles_elapsed.max,smsp__inst_executed.sum \
    --clock-control none --csv --log-file gpurun_out/balance/all.csv python tools/ncu_target.py > gpurun_out/balance/all.log 2>&1
python - <<'PY'
import csv, collections
rows = collections.OrderedDict()
for r in csv.DictReader(l for l in open("gpurun_out/balance/all.csv") if l.startswith('"')):
    d = rows.setdefault(int(r["ID"]), {"name": r["Kernel Name"].split("(")[0].replace("pb::", "")})
    d[r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
agg = collections.OrderedDict()
for d in rows.values():
    a = agg.setdefault(d["name"], [0, 0.0, 0.0, 0.0, 0.0])
    dur = d["gpu__time_duration.sum"] / 1e6
    a[0] += 1; a[1] += dur; a[2] += dur * d["sm__cycles_active.avg"] / d["sm__cycles_elapsed.max"]
    if d["smsp__inst_executed.sum"] > 5e6:       # wide launches only
        a[3] += dur; a[4] += dur * d["sm__cycles_active.avg"] / d["sm__cycles_elapsed.max"]
print(f"{'kernel':28s} launches  total_ms  sm-active-weighted_ms  idle_ms | wide launches: total  active")
for k, a in agg.items():
    print(f"{k:28s} {a[0]:6d} {a[1]:9.2f} {a[2]:12.2f} {a[1]-a[2]:14.2f} | {a[3]:9.2f} {a[4]:9.2f}")
PY
