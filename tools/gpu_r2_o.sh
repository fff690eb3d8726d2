#!/bin/bash
O=gpurun_out/r2o; mkdir -p $O
timeout -k 10 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; tail -3 $O/pytest_gpu.log
for P in 1 2 3; do timeout 600 python bench.py --parts $P --no-cpu-baseline --no-history --no-fudge-line --steps 5 --warmup 3 > $O/bench_c2_parts$P.json 2> $O/bench_parts$P.err; cut -c1-130 $O/bench_c2_parts$P.json; done
for C in 1 5; do timeout 900 python bench.py --config $C --no-cpu-baseline --no-history --no-fudge-line --steps 2 --warmup 3 > $O/bench_c$C.json 2> $O/bench_c$C.err; cut -c1-130 $O/bench_c$C.json; tail -n 2 $O/bench_c$C.err; done
