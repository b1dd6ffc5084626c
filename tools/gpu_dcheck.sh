#!/bin/bash
N=${1:-2}; TAG=${2:-r02j}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
   tools/dist_driver_check.py > gpurun_out/dist_driver_check_n${N}_$TAG.json 2> gpurun_out/dist_driver_check_n${N}_$TAG.err
echo "driver check rc=$?"; grep "^{" gpurun_out/dist_driver_check_n${N}_$TAG.json | cut -c1-2200; grep -iE "error|Traceback" -A5 gpurun_out/dist_driver_check_n${N}_$TAG.err | head -30
