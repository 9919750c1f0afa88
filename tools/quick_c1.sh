#!/bin/bash
# small-fft kernels: parity tests (RX and TX), then configs[1] and configs[0] kernel times
python -m pytest tests -m gpu -x -q -k "short_window or small_fft or fft256 or fft_len_32 or radio or c1 or hier or rolloff or single_sync or several_carrier" 2>&1 | tail -5
for c in 1 0; do
python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-agc > gpurun_out/qc$c.json 2> gpurun_out/qc$c.err || tail -5 gpurun_out/qc$c.err
python -c "
import json; d=json.loads([l for l in open('gpurun_out/qc$c.json') if l.startswith('{')][-1]); k=d['kernels_ms_per_step']; print('config $c value %.0f ms %.4f' % (d['value'], d['ms_per_step']), {a: round(b,4) for a,b in k.items() if b > 0.02})"
done
