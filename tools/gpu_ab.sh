#!/bin/bash
# A/B of kernel variants selected by environment: tools/gpu_ab.sh TAG "ENV1=a ENV2=b" "ENV1=c" ...
TAG=$1; shift
mkdir -p gpurun_out
i=0
for E in "$@"; do
i=$((i+1))
env $E timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-reference-schedule > gpurun_out/bench_ab${i}_$TAG.json 2> gpurun_out/bench_ab${i}_$TAG.err
python - <<PY
import json
t=open('gpurun_out/bench_ab${i}_$TAG.json').read()
d=json.loads(t[t.index('{'):])
print('$E ms/step %.3f'%d['ms_per_step'], 'value %.0f'%d['value'], {k: round(v,3) for k,v in d['roofline']['kernel_ms_per_step'].items() if 'multislice' in k})
PY
done
