#!/bin/bash
# 2-GPU check with the DRIVER's flags (defaults: CPU baseline on rank 0, extra lines skipped for world > 1), weak and strong; reference arm under torchrun
O=gpurun_out/r3f; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > $O/bench_c2_n2_weak.json 2> $O/n2_weak.err; echo "exit $?"; cut -c1-260 $O/bench_c2_n2_weak.json
timeout 600 $TR --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 --scaling strong > $O/bench_c2_n2_strong.json 2> $O/n2_strong.err; echo "exit $?"; cut -c1-260 $O/bench_c2_n2_strong.json
timeout 600 $TR --master-port 29514 bench.py --gpus 2 --steps 2 --warmup 3 --config 3 > $O/bench_c3_n2_weak.json 2> $O/n2_c3.err; echo "exit $?"; cut -c1-260 $O/bench_c3_n2_weak.json
