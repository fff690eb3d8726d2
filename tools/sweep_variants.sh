#!/bin/bash
# Compile-time variants (variants/*.so, built by tools/build_variants.sh with -DPB_... flags) through tools/sweep_sampler.py,
# one process each.  CFGS = the "G,T,generic" configurations every library is timed with.
for so in "" variants/*.so; do
  echo "== ${so:-default build}"
  PETITE_B200_LIB=${so:+$PWD/$so} timeout -k 10 300 python tools/sweep_sampler.py ${N:-100000} ${CFGS:-4,1,0} 2>&1 | cut -c1-700 | grep -v "^$"
done
