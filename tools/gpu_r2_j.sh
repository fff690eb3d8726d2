#!/bin/bash
# round-2 GPU call J: heavy-tile cost factor (2 / 4 / 8 / off) on the SM step and the dark passes; DarkAnn sampler form; tests
O=gpurun_out/r2j; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_probes.py tests/test_gpu_dark.py tests/test_gpu_sampling_api.py -m gpu -q 2>&1 | tail -12 > $O/pytest.log
for so in "" variants/libpb_heavy2.so variants/libpb_heavy8.so; do
  echo "== ${so:-default (heavy cost 4)}"
  PETITE_B200_LIB=${so:+$PWD/$so} SWEEP_PROFILING=1 timeout 300 python tools/sweep_sampler.py 100000 4,2,0 2>&1 | cut -c1-400
  PETITE_B200_LIB=${so:+$PWD/$so} timeout 300 python tools/dark_profile.py 3 5 2>&1 | cut -c1-330
done > $O/heavy_cost.log 2>&1
echo "== tile_norm off" >> $O/heavy_cost.log
PB_TILE_NORM=0 SWEEP_PROFILING=1 timeout 300 python tools/sweep_sampler.py 100000 4,2,0 2>&1 | cut -c1-400 >> $O/heavy_cost.log
PB_TILE_NORM=0 timeout 300 python tools/dark_profile.py 3 5 2>&1 | cut -c1-330 >> $O/heavy_cost.log
timeout 600 python tools/tile_timeline.py 100000 6 40 > $O/tile_timeline.txt 2>&1
tail -5 $O/pytest.log; cat $O/heavy_cost.log; grep "^==\|per-SM" $O/tile_timeline.txt
