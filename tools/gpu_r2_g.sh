#!/bin/bash
# round-2 GPU call G: is the dark-pass slowdown the tile normalisation?  + training schedules
O=gpurun_out/r2g; mkdir -p $O
{
for cfg in "4 0" "4 1" "8 0" "8 1"; do set -- $cfg
  echo "== G_dark=$1 T=1 tile_norm=$2"; PB_SAMPLE_G_DARK=$1 PB_SAMPLE_T_DB=1 PB_TILE_NORM=$2 timeout 300 python tools/dark_profile.py 3 5
done
echo "== prev"; PETITE_B200_LIB=$PWD/variants/libpb_prev.so timeout 300 python tools/dark_profile.py 3
} > $O/dark_profile.log 2>&1
timeout 900 python tools/exp_train.py Brem PairProd > $O/exp_train.log 2>&1
cut -c1-300 $O/dark_profile.log; cat $O/exp_train.log
