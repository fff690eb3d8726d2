#!/bin/bash
# after the final evidence run: the full GPU suite once more (tests/test_train.py changed), and the wave sizes of the captured launch
O=gpurun_out/r3z; mkdir -p $O
timeout -k 10 900 python -m pytest tests -m gpu -q --durations=5 > $O/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 $O/pytest_gpu.log
rm -f $O/waves.log; PB_GRAPH=0 PB_LOG_WAVES=$O/waves.log timeout 200 python tools/ncu_target.py --profiling 2 | cut -c1-100; sed -n 6,7p $O/waves.log
timeout 600 python tools/tile_timeline.py 100000 6 > $O/tile_timeline_w6.txt 2>&1; head -1 $O/tile_timeline_w6.txt | cut -c1-200
