#!/bin/bash
TAG=${1:-r02r}
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-reference-schedule"
timeout 900 ncu --set full --clock-control none --import-source on \
   -k regex:'k_multislice_fwd|k_multislice_bwd|k_ms3_fwd|k_ms3_bwd' -s 6 -c 2 -o gpurun_out/ms_$TAG -f $CMD > gpurun_out/ncu_ms_$TAG.log 2>&1
echo "ncu ms rc=$?"; ls -la gpurun_out/ms_$TAG.ncu-rep
