#!/bin/bash
NG=${1:-8}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $NG --steps 3 --warmup 3 > gpurun_out/bench_g$NG.json 2> gpurun_out/bench_g$NG.err
echo "== bench rc=$?"; tail -3 gpurun_out/bench_g$NG.err; python - <<PY
import json
for line in open('gpurun_out/bench_g$NG.json'):
    if line.startswith('{"metric'):
        r=json.loads(line); print(r['ms_per_step'], r['value'], 'e2e', r['e2e']['ms_per_step'], r['roofline']['achieved'], r['clocks'], r['phases_ms'])
PY
head -c 300 gpurun_out/bench_g$NG.json
