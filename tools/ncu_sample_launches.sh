#!/bin/bash
# Per-launch duration and SM-active share of every sampler launch of one config-2 batch, tile kernel vs streaming kernel.
mkdir -p gpurun_out/launch_cmp
for s in 0 1; do
  PB_SAMPLE_STREAM=$s PB_SAMPLE_TOKENS=${TOK:-8} PB_SAMPLE_G=${G:-8} timeout -k 10 400 ncu --metrics gpu__time_duration.sum,sm__cycles_active.avg,sm__cycles_elapsed.max,smsp__inst_executed.sum \
    --clock-control none -k regex:k_sample --csv --log-file gpurun_out/launch_cmp/stream$s.csv python tools/ncu_target.py > gpurun_out/launch_cmp/stream$s.log 2>&1
done
python - <<'PY'
import csv, collections
def load(p):
    rows = collections.OrderedDict()
    for r in csv.DictReader(l for l in open(p) if l.startswith('"')):
        rows.setdefault(int(r["ID"]), {})[r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
    return list(rows.values())
a, b = load("gpurun_out/launch_cmp/stream0.csv"), load("gpurun_out/launch_cmp/stream1.csv")
print("launch  inst(M)  tile_us act%   stream_us act%")
ta = tb = 0
for i, (x, y) in enumerate(zip(a, b)):
    da, db = x["gpu__time_duration.sum"] / 1e3, y["gpu__time_duration.sum"] / 1e3
    ta += da; tb += db
    print(f"{i:4d} {x['smsp__inst_executed.sum']/1e6:9.2f} {da:9.1f} {100*x['sm__cycles_active.avg']/x['sm__cycles_elapsed.max']:5.1f} {db:9.1f} {100*y['sm__cycles_active.avg']/y['sm__cycles_elapsed.max']:5.1f}")
print("total us", ta, tb)
PY
