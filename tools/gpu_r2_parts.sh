#!/bin/bash
O=gpurun_out/r2y; mkdir -p $O
for C in 2 1 3 4 5; do for P in 1 2; do
  timeout 900 python bench.py --config $C --parts $P --no-cpu-baseline --no-history --no-fudge-line --steps 3 --warmup 3 > $O/bench_c${C}_p$P.json 2> $O/bench_c${C}_p$P.err
  echo "config $C parts $P: $(cut -c1-110 $O/bench_c${C}_p$P.json)"
done; done
