#!/bin/bash
# Round-end evidence run (one GPU): full -m gpu suite, bench line, launch list of the bench command, full captures of the
# two kernels changed since r01e.
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests -m gpu -x -q --durations=5 > gpurun_out/pytest_gpu_r01f.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/pytest_gpu_r01f.log
timeout -k 10 600 python bench.py > gpurun_out/bench_r01f.json 2> gpurun_out/bench_r01f.err; echo "bench exit $?"; tail -2 gpurun_out/bench_r01f.err; cut -c1-400 gpurun_out/bench_r01f.json
timeout -k 10 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r01f_reference.json 2>/dev/null; cut -c1-300 gpurun_out/bench_r01f_reference.json
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r01f.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/launches_r01f.log 2>&1; echo "launch list exit $?"; wc -l gpurun_out/launches_r01f.csv
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
gzip -f gpurun_out/launches_r01f.csv
