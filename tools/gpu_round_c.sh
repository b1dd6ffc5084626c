#!/bin/bash
TAG=${1:-r02m}
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_$TAG.log 2>&1
echo "pytest rc=$?"; tail -12 gpurun_out/pytest_gpu_$TAG.log
bash tools/gpu_quick.sh $TAG
