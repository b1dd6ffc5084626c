#!/bin/bash
# three-pass (default) vs five-pass (ADM_MS3=0) multislice kernels: parity tests, then the headline bench for both
TAG=${1:-r04a}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/pytest_gpu_$TAG.log 2>&1
echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu_$TAG.log
for V in 1 0; do
ADM_MS3=$V timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-reference-schedule > gpurun_out/bench_ms3_${V}_$TAG.json 2> gpurun_out/bench_ms3_${V}_$TAG.err
python - <<PY
import json
t=open('gpurun_out/bench_ms3_${V}_$TAG.json').read()
d=json.loads(t[t.index('{'):])
print('ADM_MS3=$V ms/step %.3f'%d['ms_per_step'], 'value %.0f'%d['value'], {k: round(v,3) for k,v in d['roofline']['kernel_ms_per_step'].items() if 'multislice' in k})
PY
done
