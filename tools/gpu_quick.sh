#!/bin/bash
TAG=${1:-r02f}
mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-reference-schedule > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
python - <<PY
import json
t=open('gpurun_out/bench_$TAG.json').read()
d=json.loads(t[t.index('{'):])
print('N=1 ms/step %.3f'%d['ms_per_step'], 'value %.0f'%d['value'], {k: round(v,3) for k,v in d['roofline']['kernel_ms_per_step'].items()})
PY
