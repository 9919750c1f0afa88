#!/bin/bash
# ncu --set full of the small kernels between sync and the frame kernel (headline workload)
set -u
O=gpurun_out
FULL="python bench.py --steps 1 --warmup 3 --headline-only --no-cpu-baseline --no-e2e --no-agc"
ncu --set full --clock-control none --import-source on -k regex:"plateau_kernel|trig_scatter_kernel|cfo_kernel" -c 15 -f -o $O/r2_small $FULL > $O/ncu_small.log 2>&1
ls -la $O/r2_small.ncu-rep
