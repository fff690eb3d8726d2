#!/bin/bash
# final evidence run of round 2 (after the drain pause / hot_cospi / retrained-table changes) (one GPU): full -m gpu suite, smoke, bench lines (config 2 with the CPU arm, reference arm, configs 1 3 4 5),
# launch list, ncu captures of the four big kernels, tile timeline, latency
O=gpurun_out/r3z; mkdir -p $O
timeout -k 10 1500 python -m pytest tests -m gpu -q --durations=8 > $O/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 $O/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -1 $O/smoke.log | cut -c1-300
timeout -k 10 900 python bench.py > $O/bench_c2.json 2> $O/bench_c2.err; echo "bench exit $?"; cut -c1-200 $O/bench_c2.json
timeout -k 10 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_c2_reference.json 2>/dev/null; cut -c1-200 $O/bench_c2_reference.json
for C in 1 3 4 5; do timeout 1200 python bench.py --config $C --no-cpu-baseline --steps 2 --warmup 3 > $O/bench_c$C.json 2> $O/bench_c$C.err; cut -c1-150 $O/bench_c$C.json; done
PB_GRAPH=0 timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches.csv python bench.py --steps 1 --warmup 0 --parts 1 --no-cpu-baseline --no-history --no-fudge-line --no-retrained-line > $O/launches.log 2>&1; echo "launch list exit $?"; wc -l $O/launches.csv
python - <<'PY'
import csv, collections
agg = collections.OrderedDict()
for r in csv.DictReader(l for l in open("gpurun_out/r3z/launches.csv") if l.startswith('"')):
    k = r["Kernel Name"].split("(")[0].replace("pb::", "").replace("void ", "")
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += float(r["Metric Value"].replace(",", "")) / 1e6
tot = sum(a[1] for a in agg.values())
with open("gpurun_out/r3z/launches_summary.txt", "w") as f:
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"{k:40s} {a[0]:6d} launches {a[1]:10.2f} ms {100 * a[1] / tot:6.1f} %\n")
print(open("gpurun_out/r3z/launches_summary.txt").read())
PY
gzip -f $O/launches.csv
for k in k_sample k_loop k_emit k_finalize; do bash tools/ncu_one.sh r02f $k $k 20; done
timeout 600 python tools/tile_timeline.py 100000 6 22 40 > $O/tile_timeline.txt 2>&1
timeout 300 python tools/latency.py > $O/latency.json 2>/dev/null
timeout 300 python tools/dark_profile.py 3 5 > $O/dark_profile.log 2>&1
