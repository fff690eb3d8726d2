#!/bin/bash
# A/B of the L2 prefetch in k_finalize / k_emit and of drain-pause thresholds; training-weight experiment (row f-2)
O=gpurun_out/r3b; mkdir -p $O
run() { # name lib drain
  PB_DRAIN_LANES=$3 PETITE_B200_LIB=${2:+$PWD/$2} SWEEP_PROFILING=2 timeout -k 10 240 python tools/sweep_sampler.py 100000 4,2,0 2>&1 | grep -v "^$" | tail -1 | cut -c1-900 > $O/ab_$1.json
  echo "$1: $(cat $O/ab_$1.json | cut -c1-420)"
}
run pf1 "" 16
run pf0 variants/libpb_pf0.so 16
run fin7 variants/libpb_fin7.so 16
run d32 "" 32
run d28 "" 28
EXP_TRIALS=2000 timeout -k 10 600 python tools/exp_train_pow.py Brem PairProd > $O/exp_train_pow.log 2>&1; cat $O/exp_train_pow.log | cut -c1-400
