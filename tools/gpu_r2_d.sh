#!/bin/bash
# round-2 GPU call D: dark-pass sampler variants, dark tests, parts 1 vs 2
O=gpurun_out/r2d; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_dark.py tests/test_gpu_probes.py tests/test_gpu_sampling_api.py -m gpu -q 2>&1 | tail -15 > $O/pytest_dark.log
{
echo "== prev build (generic kernel, T=1)"; PETITE_B200_LIB=$PWD/variants/libpb_prev.so timeout 300 python tools/dark_profile.py 3 5
echo "== current, one generic launch"; PB_SAMPLE_SPLIT_DB=0 timeout 300 python tools/dark_profile.py 3 5
echo "== split, T_DB=1"; PB_SAMPLE_T_DB=1 timeout 300 python tools/dark_profile.py 3 5
echo "== split, T_DB=2"; PB_SAMPLE_T_DB=2 timeout 300 python tools/dark_profile.py 3 5
echo "== split, T_DB=2, G=2"; PB_SAMPLE_G=2 PB_SAMPLE_T_DB=2 timeout 300 python tools/dark_profile.py 3 5
echo "== split, T_DB=1, G=8"; PB_SAMPLE_G=8 PB_SAMPLE_T_DB=1 timeout 300 python tools/dark_profile.py 3 5
echo "== split, minb 5, T_DB=2"; PETITE_B200_LIB=$PWD/variants/libpb_db5.so PB_SAMPLE_T_DB=2 timeout 300 python tools/dark_profile.py 3 5
echo "== split, minb 5, T_DB=1"; PETITE_B200_LIB=$PWD/variants/libpb_db5.so PB_SAMPLE_T_DB=1 timeout 300 python tools/dark_profile.py 3 5
} > $O/dark_profile.log 2>&1
for P in 1 2; do timeout 600 python bench.py --parts $P --no-cpu-baseline --no-history --no-fudge-line --steps 5 --warmup 3 > $O/bench_parts$P.json 2> $O/bench_parts$P.err; done
tail -6 $O/pytest_dark.log; cut -c1-420 $O/dark_profile.log; cut -c1-120 $O/bench_parts1.json $O/bench_parts2.json
