#!/bin/bash
# ncu --set full of the two big kernels on configs[1] (fft_len 64, 4096 streams); fewer streams keep the
# replay short -- the per-sample behaviour does not depend on the stream count
set -u
O=gpurun_out
CMD="python bench.py --config 1 --streams 1024 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-agc"
$CMD > $O/c1_plain.json 2> $O/c1_plain.err || { tail -5 $O/c1_plain.err; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:sync_metric_warpn -c 1 -f -o $O/r2_sync_c1 $CMD > $O/ncu_sync_c1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rx_framew -c 1 -f -o $O/r2_frame_c1 $CMD > $O/ncu_frame_c1.log 2>&1
ls -la $O/*.ncu-rep
