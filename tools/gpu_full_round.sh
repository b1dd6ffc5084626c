#!/bin/bash
TAG=${1:-r04}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_$TAG.log 2>&1
echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu_$TAG.log
timeout 900 python bench.py --config A B C E > gpurun_out/other_configs_$TAG.jsonl 2> gpurun_out/other_configs_$TAG.err
echo "configs rc=$?"
python - <<PY
import json
for l in open('gpurun_out/other_configs_$TAG.jsonl'):
    if l.startswith('{'):
        d=json.loads(l); c=d.get('config',{}); print(c.get('workload','')[:60], '|', c.get('execution'), '| ms/step %.3f'%d['ms_per_step'], 'value %.0f'%d['value'], d.get('clocks',{}).get('sm_mhz'))
PY
( time timeout 900 python bench.py > gpurun_out/bench_default_$TAG.json 2> gpurun_out/bench_default_$TAG.err ) 2>&1 | grep real
echo "bench rc=$?"; python - <<PY
import json
t=open('gpurun_out/bench_default_$TAG.json').read()
d=json.loads(t[t.index('{'):])
print('ms/step %.3f'%d['ms_per_step'], 'value %.0f'%d['value'], 'e2e %.0f'%d['e2e']['value'], 'cpu', d['cpu_baseline'] and d['cpu_baseline']['value'], d['roofline']['frac'], d['roofline']['step_frac'], d['reference_default_schedule'])
PY
