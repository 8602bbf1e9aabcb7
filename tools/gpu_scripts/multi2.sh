#!/bin/bash
NG=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus_$NG.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 tools/dist_check.py > gpurun_out/dist_check_$NG.json 2> gpurun_out/dist_check_$NG.err
echo "== dist_check rc=$?"; tail -5 gpurun_out/dist_check_$NG.err; python -c "
import json; r=json.load(open('gpurun_out/dist_check_$NG.json')); print('ok', r['ok'], 'checks', len(r['checks']), 'failed', [c for c in r['checks'] if not c['ok']][:5], r.get('seed_agreement'))"
timeout 900 $TR --master-port 29513 tools/config_dist.py c5 c4 --scale ${2:-0.05} > gpurun_out/cfgdist_small_$NG.json 2> gpurun_out/cfgdist_small_$NG.err
echo "== config_dist rc=$?"; tail -5 gpurun_out/cfgdist_small_$NG.err; cut -c1-1500 gpurun_out/cfgdist_small_$NG.json
