#!/bin/bash
N=${1:-8}; TAG=${2:-r02z}; shift; shift
mkdir -p gpurun_out
for S in "$@"; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
   bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline --no-reference-schedule --grad-slabs $S > gpurun_out/bench_n${N}_s${S}_$TAG.json 2> gpurun_out/bench_n${N}_s${S}_$TAG.err
python - <<PY
import json
t=open('gpurun_out/bench_n${N}_s${S}_$TAG.json').read()
try:
    d=json.loads(t[t.index('{'):]); print('N=$N slabs $S', round(d['ms_per_step'],3), round(d['value']), round(d['roofline']['kernel_ms_per_step'].get('adm_p2p_reduce_adam',0),3), d['dp_check']['max_abs_replica_difference'], d['dp_check']['p2p_update_vs_allgather_sum_adam_max_abs'], d['dp_check']['update'][:70])
except Exception as e: print('N=$N slabs $S fail', e, open('gpurun_out/bench_n${N}_s${S}_$TAG.err').read()[-800:])
PY
done
