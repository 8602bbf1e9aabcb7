#!/bin/bash
NG=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus_$NG.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py > gpurun_out/dist_check_$NG.json 2> gpurun_out/dist_check_$NG.err
echo "== dist_check rc=$?"; tail -5 gpurun_out/dist_check_$NG.err; cat gpurun_out/dist_check_$NG.json | cut -c1-3000
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $NG --steps 3 --warmup 3 > gpurun_out/bench_g$NG.json 2> gpurun_out/bench_g$NG.err
echo "== bench rc=$?"; tail -5 gpurun_out/bench_g$NG.err; cat gpurun_out/bench_g$NG.json
