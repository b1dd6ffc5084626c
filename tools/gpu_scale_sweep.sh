#!/bin/bash
N=${1:-2}; TAG=${2:-r02d}
mkdir -p gpurun_out
for S in 1 2 4 8; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
   bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline --no-reference-schedule --grad-slabs $S > gpurun_out/bench_n${N}_s${S}_$TAG.json 2> gpurun_out/bench_n${N}_s${S}_$TAG.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_n${N}_s${S}_$TAG.json'))
print('slabs $S', 'ms/step', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'], {k: round(v,3) for k,v in d['roofline']['kernel_ms_per_step'].items()}, d['dp_check'])
PY
done
