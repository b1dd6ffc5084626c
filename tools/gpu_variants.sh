#!/bin/bash
TAG=${1:-r02s}
mkdir -p gpurun_out
for V in 0 2 3; do
ADM_FWD_VARIANT=$V timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-reference-schedule > gpurun_out/bench_v${V}_$TAG.json 2> gpurun_out/bench_v${V}_$TAG.err
python - <<PY
import json
t=open('gpurun_out/bench_v${V}_$TAG.json').read()
d=json.loads(t[t.index('{'):])
print('variant $V ms/step %.3f'%d['ms_per_step'], 'value %.0f'%d['value'], {k: round(v,3) for k,v in d['roofline']['kernel_ms_per_step'].items() if 'multislice' in k})
PY
done
