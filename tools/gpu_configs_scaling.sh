#!/bin/bash
# the other named shapes (A, B, C) data parallel on N ranks (run under gpurun --gpus >= N)
TAG=${1:-r03}; shift
mkdir -p gpurun_out
for N in "$@"; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
     bench.py --config A B C --gpus $N --steps 50 --warmup 5 > gpurun_out/other_configs_n${N}_$TAG.jsonl 2> gpurun_out/other_configs_n${N}_$TAG.err
  echo "N=$N rc=$?"
  python - <<PY
import json
try:
    for l in open("gpurun_out/other_configs_n${N}_$TAG.jsonl"):
        if not l.startswith('{'): continue
        d=json.loads(l); print(d["n_gpus"], d["config"]["workload"][:48], d["config"]["cuda_graph"], round(d["ms_per_step"],3), round(d["value"]))
except Exception as e:
    print('fail', e)
print(open("gpurun_out/other_configs_n${N}_$TAG.err").read()[-600:])
PY
done
