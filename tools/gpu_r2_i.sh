#!/bin/bash
O=gpurun_out/r2i; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_probes.py tests/test_gpu_showers.py tests/test_gpu_replay.py tests/test_gpu_dark.py -m gpu -q -k "not exact_decisions" 2>&1 | tail -8 > $O/pytest.log
SWEEP_PROFILING=2 timeout 300 python tools/sweep_sampler.py 100000 4,2,0 > $O/sweep_prof2.log 2>&1
SWEEP_PROFILING=0 timeout 300 python tools/sweep_sampler.py 100000 4,2,0 > $O/sweep_prof0.log 2>&1
timeout 300 python tools/dark_profile.py 3 5 > $O/dark_profile.log 2>&1
tail -4 $O/pytest.log; cut -c1-500 $O/sweep_prof2.log $O/sweep_prof0.log; cut -c1-330 $O/dark_profile.log
