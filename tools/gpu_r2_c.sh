#!/bin/bash
# round-2 GPU call C: A/B of the previous build vs the current one (per-kernel), tail check, tests, bench line
O=gpurun_out/r2c; mkdir -p $O
SWEEP_PROFILING=2 CFGS="4,2,0" tools/sweep_variants.sh > $O/sweep_prof2.log 2>&1
SWEEP_PROFILING=0 CFGS="4,2,0" tools/sweep_variants.sh > $O/sweep_prof0.log 2>&1
timeout 600 python tools/tile_timeline.py 100000 6 22 40 > $O/tile_timeline.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -25 > $O/pytest.log
timeout 600 python bench.py --no-cpu-baseline --steps 5 --warmup 3 > $O/bench_c2.json 2> $O/bench_c2.err
cut -c1-500 $O/sweep_prof2.log $O/sweep_prof0.log; grep "^==\|per-SM" $O/tile_timeline.txt; tail -5 $O/pytest.log; cut -c1-300 $O/bench_c2.json
