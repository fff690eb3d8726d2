#!/bin/bash
# 2-GPU run: weak and strong scaling of config 2, config 3 weak
O=gpurun_out/r2m; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
COMMON="bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline --no-history --no-fudge-line"
timeout 600 $TR --master-port 29511 $COMMON > $O/bench_c2_n2_weak.json 2> $O/n2_weak.err
timeout 600 $TR --master-port 29512 $COMMON --scaling strong > $O/bench_c2_n2_strong.json 2> $O/n2_strong.err
timeout 600 $TR --master-port 29513 $COMMON --config 3 > $O/bench_c3_n2_weak.json 2> $O/n2_c3.err
for f in $O/bench_c*_n2_*.json; do echo $f; cut -c1-220 $f; done; for f in $O/n2_*.err; do tail -n 2 $f | cut -c1-300; done
