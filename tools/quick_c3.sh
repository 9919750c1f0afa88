#!/bin/bash
# quick probe of BASELINE configs[3] (fft_len 2048) on the GPU box: parity tests of that path + kernel times
python -m pytest tests/test_gpu_parity.py tests/test_round2.py -m gpu -x -q -k "2048 or c4 or guard" 2>&1 | tail -3
python bench.py --config 3 --streams ${1:-256} --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/c3.json 2>gpurun_out/c3.err || tail -5 gpurun_out/c3.err
python -c "
import json; d=json.loads(open('gpurun_out/c3.json').read().strip().split(chr(10))[-1]); k=d['kernels_ms_per_step']; print('value %.0f ms %.3f chain %.3f' % (d['value'], d['ms_per_step'], d['roofline']['chain_frac']), {a: round(b, 3) for a, b in list(k.items())[:3]})"
