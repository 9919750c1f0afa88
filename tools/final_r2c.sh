#!/bin/bash
# last check of the round: whole GPU suite, smoke, ncu launch list of the headline command on the final library
set -u
O=gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -2 | tee $O/final_tests.log
python __graft_entry__.py smoke 2>&1 | tail -1 | tee $O/final_smoke.log
CMD="python bench.py --steps 3 --warmup 3 --headline-only --no-cpu-baseline --no-e2e --no-agc"
$CMD > $O/r2_prof_plain.json 2> $O/r2_prof_plain.err || tail -3 $O/r2_prof_plain.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 -k regex:"sync_metric|plateau|trig_|cfo_|rx_frame|chain_" --csv --log-file $O/r2_ncu_launches.csv $CMD > $O/r2_ncu_launches.log 2>&1
tail -c 300 $O/r2_prof_plain.json
