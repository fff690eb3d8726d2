#!/bin/bash
# tally kernel (match-aggregated), full GPU suite, default bench line (with the retrained-maps line)
O=gpurun_out/r3e; mkdir -p $O
timeout 200 python tools/time_tally.py > $O/time_tally.log 2>&1; tail -4 $O/time_tally.log
timeout -k 10 900 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 $O/pytest_gpu.log
timeout -k 10 900 python bench.py > $O/bench_c2.json 2> $O/bench_c2.err; echo "bench exit $?"; cut -c1-200 $O/bench_c2.json; tail -3 $O/bench_c2.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r3e/bench_c2.json"))
for k in ("value", "ms_per_step", "e2e", "single_stream", "maxF_fudge_4", "retrained_maps"):
    print(k, json.dumps(d.get(k))[:600])
print({k: round(v["ms_per_step"], 2) for k, v in d["kernels"].items()})
PY
