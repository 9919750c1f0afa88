#!/bin/bash
# whole GPU suite, then kernel times of the headline and configs[1], [3] (no e2e / CPU legs)
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for c in 2 1 3; do
python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-agc > gpurun_out/qa$c.json 2> gpurun_out/qa$c.err || tail -5 gpurun_out/qa$c.err
python -c "
import json; d=json.loads([l for l in open('gpurun_out/qa$c.json') if l.startswith('{')][-1]); k=d['kernels_ms_per_step']; print('config $c value %.0f ms %.4f' % (d['value'], d['ms_per_step']), {a: round(b,4) for a,b in k.items() if b > 0.015})"
done
