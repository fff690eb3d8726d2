#!/bin/bash
# final bench lines of the round: config 2 with the CPU arm, then configs 1, 3, 4, 5 (defaults)
O=gpurun_out/r2x; mkdir -p $O
timeout -k 10 900 python bench.py > $O/bench_c2.json 2> $O/bench_c2.err; cut -c1-160 $O/bench_c2.json
for C in 1 3 4 5; do timeout 1200 python bench.py --config $C --no-cpu-baseline --steps 3 --warmup 3 > $O/bench_c$C.json 2> $O/bench_c$C.err; cut -c1-160 $O/bench_c$C.json; done
