#!/bin/bash
TAG=${1:-r04g}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/pytest_gpu_$TAG.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu_$TAG.log
ADM_MS_VERBOSE=1 bash tools/gpu_ab.sh $TAG "ADM_MS3_FWD=0" "ADM_MS3_FWD=1" "ADM_MS3_FWD=2" "ADM_MS3_FWD=0"
grep -h "adm\]" gpurun_out/bench_ab1_$TAG.err | sort | uniq -c
