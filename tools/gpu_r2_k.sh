#!/bin/bash
# round-2 GPU call K: config-4 ensemble test, DarkAnn golden bound, bench lines of configs 1, 3, 4, 5, dark ncu capture
O=gpurun_out/r2k; mkdir -p $O
timeout 900 python -m pytest "tests/test_gpu_ensemble.py::test_observables_1e5_showers_vs_reference_and_oracle[c4_beamdump_lead_dark]" tests/test_gpu_probes.py -m gpu -q -s 2>&1 | tail -30 > $O/pytest.log
for C in 1 3 4 5; do
  timeout 900 python bench.py --config $C --no-cpu-baseline --steps 3 --warmup 3 > $O/bench_c$C.json 2> $O/bench_c$C.err
done
timeout 300 python tools/dark_profile.py 3 5 > $O/dark_profile.log 2>&1
grep -E "passed|failed|c4_beamdump" $O/pytest.log | cut -c1-900; for C in 1 3 4 5; do cut -c1-170 $O/bench_c$C.json; tail -2 $O/bench_c$C.err; done; cut -c1-300 $O/dark_profile.log
