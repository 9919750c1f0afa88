#!/bin/bash
# Round-2 end-of-round refresh on one GPU: whole GPU suite, default bench (all configurations), reference arm, smoke
set -u
O=gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -3 | tee $O/final_tests.log
python bench.py > $O/r2_final_bench.json 2> $O/r2_final_bench.err || tail -5 $O/r2_final_bench.err
python bench.py --impl reference > $O/r2_final_ref.json 2> $O/r2_final_ref.err || tail -5 $O/r2_final_ref.err
python __graft_entry__.py smoke 2>&1 | tail -2 | tee $O/final_smoke.log
python tools/bench_stage.py --stage agc2 > $O/stage_agc2.json 2>&1 || tail -3 $O/stage_agc2.json
