#!/bin/bash
# fused peer-memory update with / without NVLS multicast at N ranks
N=${1:-4}; TAG=${2:-r02y}
mkdir -p gpurun_out
for MC in 0 1; do
ADM_P2P_MULTICAST=$MC timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
   bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline --no-reference-schedule --dp p2p > gpurun_out/bench_n${N}_mc${MC}_$TAG.json 2> gpurun_out/bench_n${N}_mc${MC}_$TAG.err
python - <<PY
import json
t=open('gpurun_out/bench_n${N}_mc${MC}_$TAG.json').read()
try:
    d=json.loads(t[t.index('{'):])
    print('mc=$MC', 'ms/step %.3f'%d['ms_per_step'], 'value %.0f'%d['value'], round(d['roofline']['kernel_ms_per_step'].get('adm_p2p_reduce_adam',0),3), d['dp_check']['update'], d['dp_check']['max_abs_replica_difference'])
except Exception as e:
    print('mc=$MC failed', e); print(open('gpurun_out/bench_n${N}_mc${MC}_$TAG.err').read()[-2000:])
PY
done
