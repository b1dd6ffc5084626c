#!/bin/bash
# one gpurun round: hand-off tests first (short timeout: a dead-locked persistent kernel must not eat the budget),
# then the whole GPU suite, a bench line and the ncu launch list of a short bench.
mkdir -p gpurun_out
TAG=${1:-r02a}
timeout 600 python -m pytest tests -m gpu -x -q -k "handoff" > gpurun_out/pytest_handoff_$TAG.log 2>&1
echo "handoff rc=$?" | tee -a gpurun_out/pytest_handoff_$TAG.log
tail -3 gpurun_out/pytest_handoff_$TAG.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_$TAG.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu_$TAG.log
tail -15 gpurun_out/pytest_gpu_$TAG.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
echo "bench rc=$?"; cat gpurun_out/bench_$TAG.json; tail -3 gpurun_out/bench_$TAG.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-reference-schedule > gpurun_out/ncu_ll_$TAG.log 2>&1
echo "ncu rc=$?"
