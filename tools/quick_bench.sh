#!/bin/bash
# quick headline probe on the GPU box: frame-kernel parity tests + kernel times (no e2e / CPU legs)
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "c3 or warp_frame or frame_kernel_variants or large_roundtrip" 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 --headline-only --no-cpu-baseline --no-e2e > gpurun_out/quick.json 2> gpurun_out/quick.err || tail -5 gpurun_out/quick.err
python -c "
import json; d=json.load(open('gpurun_out/quick.json')); k=d['kernels_ms_per_step']; print('value %.0f ms %.4f frame %.4f sync %.4f' % (d['value'], d['ms_per_step'], k.get('rx_framew_kernel',0), k.get('sync_metric_warp_kernel',0)))"
