#!/bin/bash
# round-2 GPU call A: full GPU test suite, 400 GeV artefacts (copied back through gpurun_out/), all-config bench lines, latency
O=gpurun_out/r2a; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -40 > $O/pytest.log
timeout 600 python tools/make_400GeV.py > $O/make_400GeV.log 2>&1
mkdir -p $O/data_400GeV; cp data_400GeV/sm_maps.npz data_400GeV/sm_maxF.npz data_400GeV/dark_maps_mV0.01.npz data_400GeV/dark_maxF.npz data_400GeV/dark_setup_*.npz $O/data_400GeV/ 2>> $O/make_400GeV.log
for C in 1 3 4 5; do
  timeout 900 python bench.py --config $C --no-cpu-baseline --steps 2 --warmup 3 > $O/bench_c$C.json 2> $O/bench_c$C.err
done
PB_GRAPH=1 timeout 300 python tools/latency.py > $O/latency_graph.json 2> $O/latency.err
PB_GRAPH=0 timeout 300 python tools/latency.py > $O/latency_stream.json 2>> $O/latency.err
tail -8 $O/pytest.log; tail -3 $O/make_400GeV.log; ls -la $O $O/data_400GeV
