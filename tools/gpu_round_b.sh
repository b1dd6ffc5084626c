#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r02b}
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_$TAG.log 2>&1
echo "pytest rc=$?"; tail -12 gpurun_out/pytest_gpu_$TAG.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
echo "bench rc=$?"; cat gpurun_out/bench_$TAG.json; tail -3 gpurun_out/bench_$TAG.err
timeout 900 python bench.py --config A B C E --steps 30 --warmup 5 > gpurun_out/bench_configs_$TAG.jsonl 2> gpurun_out/bench_configs_$TAG.err
echo "configs rc=$?"; cut -c1-400 gpurun_out/bench_configs_$TAG.jsonl; tail -3 gpurun_out/bench_configs_$TAG.err
bash tools/sanitize.sh $TAG
