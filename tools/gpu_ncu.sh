#!/bin/bash
# one ncu --set full capture of every kernel of the config-D step (1 launch each, after warm-up), + launch list
TAG=${1:-r02o}
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-reference-schedule"
timeout 900 ncu --set full --clock-control none --import-source on \
   -k regex:'k_ms3_fwd|k_ms3_bwd|k_multislice_fwd|k_multislice_bwd' -s 6 -c 2 -o gpurun_out/ms_$TAG -f $CMD > gpurun_out/ncu_ms_$TAG.log 2>&1
echo "ncu ms rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on \
   -k regex:'k_object_to_transmission|k_transmission_to_object_adj|k_reg_l1tv|k_adam|k_detector_loss' -s 15 -c 6 -o gpurun_out/small_$TAG -f $CMD > gpurun_out/ncu_small_$TAG.log 2>&1
echo "ncu small rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_ll_$TAG.log 2>&1
echo "ncu ll rc=$?"; ls -la gpurun_out/*_$TAG*
