#!/bin/bash
# compute-sanitizer memcheck + racecheck for the kernels added late in round 2: affine resampling (tilt / hologram alignment),
# batched large-frame patch steps + far field + complex frame shift (probes of 128 pixels), tiled MultiDist, exit-wave shift,
# and the final multislice kernels (paired table loads) on the 3-D edge case.
mkdir -p gpurun_out
TAG=${1:-r02z}
CASE='ptycho3d_p72_tilt or ptycho2d_p128 or ptycho3d_p128_near or fullfield64_prjoffset or ptycho3d_p72_edge or multidist128_affine or multidist144_tiled'
for tool in memcheck racecheck; do
  timeout 1500 compute-sanitizer --tool $tool --print-limit 20 \
      python -m pytest tests/test_gpu_parity.py tests/test_gpu_multidist.py -x -q -k "$CASE" > gpurun_out/sanitizer_${tool}_new_$TAG.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/sanitizer_${tool}_new_$TAG.log | tail -4
done
