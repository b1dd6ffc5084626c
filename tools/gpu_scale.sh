#!/bin/bash
# bench at N GPUs of one box (N = $1), our arm only
N=${1:-2}; TAG=${2:-r02c}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
   bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline --no-reference-schedule > gpurun_out/bench_n${N}_$TAG.json 2> gpurun_out/bench_n${N}_$TAG.err
echo "rc=$?"; cat gpurun_out/bench_n${N}_$TAG.json | cut -c1-3000; tail -5 gpurun_out/bench_n${N}_$TAG.err
