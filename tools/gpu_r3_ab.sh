#!/bin/bash
# (historical record of how profiles/r02_final/ab_*.json were made; the nc0 variant needs the node-cache option that existed up to commit 39d01aa)
# A/B of the k_loop changes (node cache, cp.async.cg, drain pause) and hot_cospi: full GPU suite on the default build, then
# tools/sweep_sampler.py (digest of 2 000 showers + config-2 timing, per kernel and in graph mode) per variant, one process each
O=gpurun_out/r3a; mkdir -p $O
timeout -k 10 900 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 $O/pytest_gpu.log
run() { # name lib drain
  PB_DRAIN_LANES=$3 PETITE_B200_LIB=${2:+$PWD/$2} SWEEP_PROFILING=2 timeout -k 10 240 python tools/sweep_sampler.py 100000 4,2,0 2>&1 | grep -v "^$" | tail -1 | cut -c1-900 > $O/ab_$1.json
  echo "$1: $(cat $O/ab_$1.json | cut -c1-420)"
}
run old variants/libpb_old.so 0
run d0 "" 0
run d16 "" 16
run d10 "" 10
run d22 "" 22
run nc0 variants/libpb_nc0.so 16
run cg0 variants/libpb_cg0.so 16
run cospi0 variants/libpb_cospi0.so 16
