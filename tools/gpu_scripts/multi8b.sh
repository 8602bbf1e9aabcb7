#!/bin/bash
NG=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29512 bench.py --gpus $NG --steps 3 --warmup 3 > gpurun_out/bench_g$NG.json 2> gpurun_out/bench_g$NG.err
echo "== bench rc=$?"; tail -2 gpurun_out/bench_g$NG.err; cut -c1-300 gpurun_out/bench_g$NG.json
timeout 900 $TR --master-port 29513 tools/config_dist.py c4 --reps 2 > gpurun_out/cfgdist_c4_$NG.json 2> gpurun_out/cfgdist_c4_$NG.err
echo "== config_dist rc=$?"; tail -3 gpurun_out/cfgdist_c4_$NG.err; cut -c1-900 gpurun_out/cfgdist_c4_$NG.json
