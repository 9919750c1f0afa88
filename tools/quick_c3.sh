#!/bin/bash
# fft_len 2048 probe on the GPU box: sync / frame parity tests, then configs[3] kernel times
python -m pytest tests/test_round2.py tests/test_gpu_parity.py -m gpu -x -q -k "short_window or fft2048 or sync_kernel_variants or c4" 2>&1 | tail -5
python bench.py --config 3 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-agc > gpurun_out/qc3.json 2> gpurun_out/qc3.err || tail -5 gpurun_out/qc3.err
python -c "
import json; d=json.loads([l for l in open('gpurun_out/qc3.json') if l.startswith('{')][-1]); k=d['kernels_ms_per_step']; print('config 3 value %.0f ms %.4f' % (d['value'], d['ms_per_step']), {a: round(b,4) for a,b in k.items() if b > 0.05})"
