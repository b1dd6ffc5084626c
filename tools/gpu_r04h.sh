#!/bin/bash
TAG=${1:-r04h}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_$TAG.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu_$TAG.log
bash tools/gpu_ab.sh $TAG "ADM_MS3=1"
for V in 1 0; do
ADM_MS3=$V timeout 600 python bench.py --config A --steps 200 --warmup 10 > gpurun_out/other_A_ms3_${V}_$TAG.jsonl 2> gpurun_out/other_A_ms3_${V}_$TAG.err
echo "config A ADM_MS3=$V rc=$?"
python - <<PY
import json
for l in open('gpurun_out/other_A_ms3_${V}_$TAG.jsonl'):
    if l.startswith('{'):
        d=json.loads(l); print({k: d.get(k) for k in ('value','ms_per_step')}, d.get('config',{}).get('execution'), d.get('graph'))
PY
done
