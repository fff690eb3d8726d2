#!/bin/bash
# round-2 GPU call F: dark sampler with the early-return integrand restored: G / T / split / sqrt variants
O=gpurun_out/r2f; mkdir -p $O
{
for cfg in "8 1 1" "8 2 1" "4 1 1" "4 2 1" "8 1 0" "16 1 1"; do set -- $cfg
  echo "== G_dark=$1 T_DB=$2 split=$3"; PB_SAMPLE_G_DARK=$1 PB_SAMPLE_T_DB=$2 PB_SAMPLE_T_DARK=1 PB_SAMPLE_SPLIT_DB=$3 timeout 300 python tools/dark_profile.py 3 5
done
echo "== IEEE sqrt, G_dark=8 T_DB=1 split"; PETITE_B200_LIB=$PWD/variants/libpb_dbsqrt.so PB_SAMPLE_G_DARK=8 PB_SAMPLE_T_DB=1 timeout 300 python tools/dark_profile.py 3 5
} > $O/dark_profile.log 2>&1
timeout 900 python -m pytest tests/test_gpu_dark.py tests/test_gpu_probes.py tests/test_gpu_sampling_api.py -m gpu -q 2>&1 | tail -8 > $O/pytest_dark.log
cut -c1-300 $O/dark_profile.log; tail -4 $O/pytest_dark.log
