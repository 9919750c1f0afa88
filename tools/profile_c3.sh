#!/bin/bash
# ncu --set full of the two big kernels on configs[3] (fft_len 2048); 128 streams keep the replay short
set -u
O=gpurun_out
CMD="python bench.py --config 3 --streams 128 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-agc"
$CMD > $O/c3_plain.json 2> $O/c3_plain.err || { tail -5 $O/c3_plain.err; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:rx_framep -c 1 -f -o $O/r2_framep3 $CMD > $O/ncu_framep3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sync_metric_warp -c 1 -f -o $O/r2_sync_c3 $CMD > $O/ncu_sync_c3.log 2>&1
ls -la $O/r2_framep3.ncu-rep $O/r2_sync_c3.ncu-rep
