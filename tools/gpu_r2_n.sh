#!/bin/bash
# round-2 GPU call N: bounded sub-steps per launch (carry-over): correctness under a small cap, timing vs cap
O=gpurun_out/r2n; mkdir -p $O
PB_LOOP_CAP=8 timeout 900 python -m pytest tests/test_gpu_showers.py tests/test_gpu_dark.py tests/test_gpu_sampling_api.py -m gpu -q 2>&1 | tail -8 > $O/pytest_cap8.log
timeout 600 python -m pytest tests/test_gpu_showers.py -m gpu -q 2>&1 | tail -4 > $O/pytest_cap0.log
for cap in 0 16 32 64 128; do
  echo "== PB_LOOP_CAP=$cap"
  PB_LOOP_CAP=$cap SWEEP_PROFILING=1 timeout 300 python tools/sweep_sampler.py 100000 4,2,0 2>&1 | cut -c1-420
  PB_LOOP_CAP=$cap SWEEP_PROFILING=0 timeout 300 python tools/sweep_sampler.py 100000 4,2,0 2>&1 | cut -c1-420
  PB_LOOP_CAP=$cap timeout 300 python tools/latency.py 2>&1 | cut -c1-900
done > $O/loop_cap.log 2>&1
tail -4 $O/pytest_cap8.log; tail -2 $O/pytest_cap0.log; cat $O/loop_cap.log
