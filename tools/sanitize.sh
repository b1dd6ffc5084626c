#!/bin/bash
# compute-sanitizer evidence for the kernels that hand shared memory to the async proxy (TMA loads / bulk copies /
# TMA reduce) and that pass wavefields between CTAs: racecheck + memcheck on small parity cases.
mkdir -p gpurun_out
TAG=${1:-r02}
CASE='test_loss_and_gradients_match_oracle_and_reference and (ptycho3d_p72_edge or ptycho2d_realimag_p64)'
for tool in memcheck racecheck; do
  timeout 1500 compute-sanitizer --tool $tool --print-limit 20 \
      python -m pytest tests/test_gpu_parity.py -x -q -k "$CASE" > gpurun_out/sanitizer_${tool}_$TAG.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/sanitizer_${tool}_$TAG.log | tail -4
done
# hand-off path (wavefields cut between CTAs) + large-frame + probe-shift kernels under memcheck
ADM_MS_GRID=3 timeout 1200 compute-sanitizer --tool memcheck --print-limit 20 \
    python -m pytest tests/test_gpu_parity.py tests/test_gpu_multidist.py -x -q -k "handoff and edge-3 or multidist128" \
    > gpurun_out/sanitizer_memcheck_handoff_$TAG.log 2>&1
echo "memcheck handoff rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitizer_memcheck_handoff_$TAG.log | tail -3
