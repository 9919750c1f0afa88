#!/bin/bash
# Round-2 final evidence on one GPU: whole GPU suite, default bench (all configurations), reference arm, smoke,
# ncu launch list of the headline command, ncu --set full of the kernels that changed late in the round.
set -u
O=gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -3 | tee $O/final_tests.log
python bench.py > $O/r2_final_bench.json 2> $O/r2_final_bench.err || tail -5 $O/r2_final_bench.err
python bench.py --impl reference > $O/r2_final_ref.json 2> $O/r2_final_ref.err || tail -5 $O/r2_final_ref.err
python __graft_entry__.py smoke 2>&1 | tail -2 | tee $O/final_smoke.log
CMD="python bench.py --steps 3 --warmup 3 --headline-only --no-cpu-baseline --no-e2e --no-agc"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 -k regex:"sync_metric|plateau|trig_|cfo_|rx_frame|chain_" --csv --log-file $O/r2_ncu_launches.csv $CMD > $O/r2_ncu_launches.log 2>&1
C1="python bench.py --config 1 --streams 1024 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-agc"
ncu --set full --clock-control none --import-source on -k regex:sync_metric_warpn -c 1 -f -o $O/r2_sync_c1 $C1 > $O/ncu_sync_c1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rx_framew -c 1 -f -o $O/r2_frame_c1 $C1 > $O/ncu_frame_c1.log 2>&1
C3="python bench.py --config 3 --streams 128 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-agc"
ncu --set full --clock-control none --import-source on -k regex:rx_framep -c 1 -f -o $O/r2_framep3 $C3 > $O/ncu_framep3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sync_metric_warp -c 1 -f -o $O/r2_sync_c3 $C3 > $O/ncu_sync_c3.log 2>&1
ls -la $O/*.ncu-rep | tail -6
