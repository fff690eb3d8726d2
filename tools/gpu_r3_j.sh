#!/bin/bash
# long-lived decays (row f-5): full GPU suite, then the config-2 timing (k_emit must not regress)
O=gpurun_out/r3j; mkdir -p $O
timeout -k 10 900 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -5 $O/pytest_gpu.log
SWEEP_PROFILING=2 timeout -k 10 240 python tools/sweep_sampler.py 100000 4,2,0 2>&1 | grep -v "^$" | tail -1 | cut -c1-900 > $O/ab_longlived.json; cut -c1-420 $O/ab_longlived.json
