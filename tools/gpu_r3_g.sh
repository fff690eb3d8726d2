#!/bin/bash
# loop-cap A/B with the drain pause in place
O=gpurun_out/r3g; mkdir -p $O
for cap in 16 64 128; do
  PB_LOOP_CAP=$cap SWEEP_PROFILING=2 timeout -k 10 240 python tools/sweep_sampler.py 100000 4,2,0 2>&1 | grep -v "^$" | tail -1 | cut -c1-900 > $O/ab_cap$cap.json
  echo "cap $cap: $(cut -c1-330 $O/ab_cap$cap.json)"
done
