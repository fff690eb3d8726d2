#!/bin/bash
O=gpurun_out/r2s; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_probes.py tests/test_gpu_showers.py tests/test_gpu_replay.py tests/test_gpu_dark.py -m gpu -q 2>&1 | tail -8 > $O/pytest.log
SWEEP_PROFILING=2 timeout 300 python tools/sweep_sampler.py 100000 4,2,0 2>&1 | cut -c1-500 > $O/sweep_prof2.log
SWEEP_PROFILING=0 timeout 300 python tools/sweep_sampler.py 100000 4,2,0 2>&1 | cut -c1-500 > $O/sweep_prof0.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-500 > $O/smoke.log
tail -5 $O/pytest.log; cat $O/sweep_prof2.log $O/sweep_prof0.log $O/smoke.log
