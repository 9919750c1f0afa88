#!/bin/bash
# DRAM traffic of the pair-of-warps kernel on configs[3] (--streams 128): every launch of the run, the full-size ones are the large ones
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:rx_framep -c 6 python bench.py --config 3 --streams 128 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-agc 2>&1 | grep -E "dram__bytes|gpu__time"
