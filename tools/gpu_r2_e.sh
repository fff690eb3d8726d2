#!/bin/bash
# round-2 GPU call E: f-3 on the GPU, dark sampler after the relative tile normalisation, SM step check
O=gpurun_out/r2e; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_dark_setup.py tests/test_gpu_dark.py -m gpu -q -s 2>&1 | tail -25 > $O/pytest_f3.log
{
echo "== current, one generic launch"; PB_SAMPLE_SPLIT_DB=0 timeout 300 python tools/dark_profile.py 3 5
echo "== split, T_DB=1"; PB_SAMPLE_T_DB=1 timeout 300 python tools/dark_profile.py 3 5
echo "== split, T_DB=2"; PB_SAMPLE_T_DB=2 timeout 300 python tools/dark_profile.py 3 5
echo "== split, T_DB=1, G=8"; PB_SAMPLE_G=8 PB_SAMPLE_T_DB=1 timeout 300 python tools/dark_profile.py 3 5
echo "== split, T_DB=2, G=8"; PB_SAMPLE_G=8 PB_SAMPLE_T_DB=2 timeout 300 python tools/dark_profile.py 3 5
} > $O/dark_profile.log 2>&1
SWEEP_PROFILING=0 timeout 300 python tools/sweep_sampler.py 100000 4,2,0 > $O/sweep_prof0.log 2>&1
tail -12 $O/pytest_f3.log; cut -c1-330 $O/dark_profile.log; cut -c1-300 $O/sweep_prof0.log
