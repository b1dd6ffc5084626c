#!/bin/bash
# weak scaling of the default bench line at N ranks (run under gpurun --gpus N)
N=${1:-8}; TAG=${2:-r02z}
mkdir -p gpurun_out
if [ "$N" = "1" ]; then
  timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 > gpurun_out/bench_n1_$TAG.json 2> gpurun_out/bench_n1_$TAG.err
else
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
     bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline --no-reference-schedule > gpurun_out/bench_n${N}_$TAG.json 2> gpurun_out/bench_n${N}_$TAG.err
fi
echo "rc=$?"
python - <<PY
import json
t=open('gpurun_out/bench_n${N}_$TAG.json').read()
try:
    d=json.loads(t[t.index('{'):])
    print('N=$N', 'ms/step %.3f'%d['ms_per_step'], 'value %.0f'%d['value'], 'e2e %.0f'%d['e2e']['value'], {k: round(v,3) for k,v in d['roofline']['kernel_ms_per_step'].items()}, d['dp_check'], d['clocks'])
except Exception as e:
    print('failed', e); print(open('gpurun_out/bench_n${N}_$TAG.err').read()[-3000:])
PY
