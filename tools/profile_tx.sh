#!/bin/bash
# ncu --set full of the warp-per-packet TX kernel, rectangular prefix and rolloff 18 (fft_len 1024 plan)
set -u
O=gpurun_out
python tools/bench_stage.py --stage tx --frames 32768 --steps 3 > $O/tx_rect.json 2>&1
python tools/bench_stage.py --stage tx --frames 32768 --steps 3 --rolloff 18 > $O/tx_roll.json 2>&1
tail -1 $O/tx_rect.json | cut -c1-200; tail -1 $O/tx_roll.json | cut -c1-200
ncu --set full --clock-control none --import-source on -k regex:tx_framew -c 1 -s 2 -f -o $O/r2_tx_rect python tools/bench_stage.py --stage tx --frames 32768 --steps 2 > $O/ncu_tx_rect.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tx_framew -c 1 -s 2 -f -o $O/r2_tx_roll python tools/bench_stage.py --stage tx --frames 32768 --steps 2 --rolloff 18 > $O/ncu_tx_roll.log 2>&1
ls -la $O/r2_tx_*.ncu-rep
