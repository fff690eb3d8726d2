#!/bin/bash
# Compile-time variants (variants/*.so, built with -DPB_TILE=... etc.) through tools/sweep_sampler.py, one process each.
for so in "" variants/*.so; do
  echo "== ${so:-default build}"
  PETITE_B200_LIB=${so:+$PWD/$so} timeout -k 10 200 python tools/sweep_sampler.py 100000 ${CFGS:-0,4,0,0} 2>&1 | cut -c1-600 | grep -v "^$"
done
