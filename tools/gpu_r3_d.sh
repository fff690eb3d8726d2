#!/bin/bash
# row f-2 with the sampler-oriented training weight: GPU training tests, the 400 GeV table set retrained (config 4), its parity tests and bench line
O=gpurun_out/r3d; mkdir -p $O/data_400GeV
timeout -k 10 600 python -m pytest tests/test_train.py -m gpu -q -s > $O/pytest_train.log 2>&1; echo "train tests exit $?"; grep -i "efficiency\|passed\|failed" $O/pytest_train.log | cut -c1-200
SWEEP_PROFILING=2 timeout -k 10 240 python tools/sweep_sampler.py 100000 4,2,0 2>&1 | grep -v "^$" | tail -1 | cut -c1-900 > $O/ab_default.json; cut -c1-420 $O/ab_default.json
timeout -k 10 300 python bench.py --config 4 --no-cpu-baseline --steps 2 --warmup 3 > $O/bench_c4_before.json 2> $O/bench_c4_before.err; cut -c1-200 $O/bench_c4_before.json
timeout -k 10 600 python tools/make_400GeV.py > $O/make_400GeV.log 2>&1; echo "make_400GeV exit $?"; tail -4 $O/make_400GeV.log | cut -c1-400
cp data_400GeV/sm_maps.npz data_400GeV/sm_maxF.npz data_400GeV/dark_maps_mV0.01.npz data_400GeV/dark_maxF.npz data_400GeV/dark_setup_lead_mV0.01.npz $O/data_400GeV/
timeout -k 10 900 python -m pytest tests -m gpu -q -k "c4_beamdump" > $O/pytest_c4.log 2>&1; echo "c4 tests exit $?"; tail -3 $O/pytest_c4.log
timeout -k 10 300 python bench.py --config 4 --no-cpu-baseline --steps 2 --warmup 3 > $O/bench_c4.json 2> $O/bench_c4.err; cut -c1-200 $O/bench_c4.json
