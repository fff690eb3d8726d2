#!/bin/bash
# round-2 GPU call B: tests after the look-up / packing changes, per-kernel step times, k_sample tile timeline
O=gpurun_out/r2b; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -25 > $O/pytest.log
SWEEP_PROFILING=2 timeout 300 python tools/sweep_sampler.py 100000 4,2,0 > $O/sweep_prof2.log 2>&1
SWEEP_PROFILING=0 timeout 300 python tools/sweep_sampler.py 100000 4,2,0 > $O/sweep_prof0.log 2>&1
timeout 600 python tools/tile_timeline.py 100000 6 12 22 40 > $O/tile_timeline.txt 2>&1
tail -5 $O/pytest.log; cat $O/sweep_prof2.log $O/sweep_prof0.log | cut -c1-600; head -50 $O/tile_timeline.txt
