#!/bin/bash
O=gpurun_out/r2p; mkdir -p $O
SWEEP_PROFILING=2 CFGS="4,2,0" tools/sweep_variants.sh > $O/sweep_prof2.log 2>&1
SWEEP_PROFILING=0 CFGS="4,2,0" tools/sweep_variants.sh > $O/sweep_prof0.log 2>&1
for P in 1 2; do timeout 600 python bench.py --parts $P --no-cpu-baseline --no-history --no-fudge-line --steps 5 --warmup 3 > $O/bench_c2_parts$P.json 2> $O/bench_parts$P.err; cut -c1-130 $O/bench_c2_parts$P.json; done
cut -c1-500 $O/sweep_prof2.log $O/sweep_prof0.log
