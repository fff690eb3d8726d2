#!/bin/bash
# variants/libpb_<name>.so for every "name:flags" argument, e.g.  tools/build_variants.sh "ph7:-DPB_PHILOX_ROUNDS=7" "t2b4:-DPB_SAMPLE_MINB_SM_T=4"
mkdir -p variants
for spec in "$@"; do
  name=${spec%%:*}; flags=${spec#*:}
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared $flags -o variants/libpb_$name.so petite_b200/csrc/engine.cu &
done
wait
ls -la variants
