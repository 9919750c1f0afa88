#!/bin/bash
# Round-2 evidence run (on the GPU box): plain bench first, then the ncu launch list of the same command, then
# ncu --set full of the two dominant kernels.  Outputs under gpurun_out/ (summaries are copied to profiles/).
set -u
O=gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --headline-only --no-cpu-baseline --no-e2e --no-agc"
$CMD > $O/r2_prof_plain.json 2> $O/r2_prof_plain.err || { tail -5 $O/r2_prof_plain.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_ncu_launches.csv $CMD > $O/r2_ncu_launches.log 2>&1
FULL="python bench.py --frames 32768 --steps 1 --warmup 3 --headline-only --no-cpu-baseline --no-e2e --no-agc"
ncu --set full --clock-control none --import-source on -k regex:rx_framew -c 1 -f -o $O/r2_frame $FULL > $O/r2_ncu_frame.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sync_metric_warp -c 1 -f -o $O/r2_sync $FULL > $O/r2_ncu_sync.log 2>&1
ls -la $O/*.ncu-rep
