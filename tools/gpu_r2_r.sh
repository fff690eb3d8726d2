#!/bin/bash
for k in k_finalize k_loop; do bash tools/ncu_one.sh r02c $k $k 20; done
