#!/bin/bash
N=${1:-2}; TAG=${2:-r02e}
mkdir -p gpurun_out
for MODE in p2p nccl; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
   bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline --no-reference-schedule --dp $MODE > gpurun_out/bench_n${N}_${MODE}_$TAG.json 2> gpurun_out/bench_n${N}_${MODE}_$TAG.err
echo "rc=$?"
python - <<PY
import json
t=open('gpurun_out/bench_n${N}_${MODE}_$TAG.json').read()
try:
    d=json.loads(t[t.index('{'):])
    print('$MODE', 'ms/step %.3f'%d['ms_per_step'], 'value %.0f'%d['value'], 'e2e %.0f'%d['e2e']['value'], {k: round(v,3) for k,v in d['roofline']['kernel_ms_per_step'].items()}, d['dp_check'])
except Exception as e:
    print('$MODE failed', e); print(open('gpurun_out/bench_n${N}_${MODE}_$TAG.err').read()[-3000:])
PY
done
