#!/bin/bash
O=gpurun_out/r3k; mkdir -p $O
run() { PETITE_B200_LIB=${2:+$PWD/$2} SWEEP_PROFILING=2 timeout -k 10 240 python tools/sweep_sampler.py 100000 4,2,0 2>&1 | grep -v "^$" | tail -1 | cut -c1-900 > $O/ab_$1.json; echo "$1: $(cut -c1-420 $O/ab_$1.json)"; }
run emcs variants/libpb_emcs.so
run spec variants/libpb_spec.so
