#!/bin/bash
O=gpurun_out/r2l; mkdir -p $O
timeout 1200 python -m pytest tests/test_gpu_ensemble.py tests/test_gpu_dark.py tests/test_gpu_sampling_api.py -m gpu -q -s 2>&1 | grep -E "passed|failed|vs reference|Error|error" | cut -c1-1500 > $O/pytest.log
for C in 4 5; do
  timeout 1200 python bench.py --config $C --no-cpu-baseline --steps 2 --warmup 3 > $O/bench_c$C.json 2> $O/bench_c$C.err
done
cat $O/pytest.log | cut -c1-400; for C in 4 5; do cut -c1-170 $O/bench_c$C.json; tail -2 $O/bench_c$C.err; done
