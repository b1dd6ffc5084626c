#!/bin/bash
# compute-sanitizer memcheck + racecheck of the three-pass multislice kernels (ms3_kernels.cuh): P = 72 (overhanging patches,
# exit wave, sparse per-gap tables, hand-off between CTAs) and P = 64 (full field, explicit patch stacks, far field).
mkdir -p gpurun_out
TAG=${1:-r04s}
CASE='(oracle_and_reference and (ptycho3d_p72_edge or ptycho3d_p72_exitwave or sparse3_p72 or fullfield64)) or (handoff and ptycho3d_p72_edge) or three_pass_gradients'
for tool in memcheck racecheck; do
  timeout 1500 compute-sanitizer --tool $tool --print-limit 20 \
      python -m pytest tests/test_gpu_parity.py -x -q -k "$CASE" > gpurun_out/sanitizer_${tool}_ms3_$TAG.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/sanitizer_${tool}_ms3_$TAG.log | tail -4
done
