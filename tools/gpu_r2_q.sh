#!/bin/bash
O=gpurun_out/r2q; mkdir -p $O
timeout -k 10 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; tail -3 $O/pytest_gpu.log
SWEEP_PROFILING=2 timeout 300 python tools/sweep_sampler.py 100000 4,2,0 2>&1 | cut -c1-500
for P in 1 2; do timeout 600 python bench.py --parts $P --no-cpu-baseline --no-history --no-fudge-line --steps 5 --warmup 3 > $O/bench_c2_parts$P.json 2> $O/bench_parts$P.err; cut -c1-130 $O/bench_c2_parts$P.json; done
