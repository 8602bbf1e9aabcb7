#!/bin/bash
# first GPU pass: kernels, parity, a small and a full bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
(lscpu | head -20; nproc; free -g | head -2) > gpurun_out/cpu.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "not tensor_core and not sweep" 2>&1 | tail -80 > gpurun_out/t_kernels.log
echo "== kernels"; tail -30 gpurun_out/t_kernels.log
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "tensor_core or sweep" 2>&1 | tail -80 > gpurun_out/t_tc.log
echo "== tc"; tail -40 gpurun_out/t_tc.log
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu 2>&1 | tail -120 > gpurun_out/t_parity.log
echo "== parity"; tail -60 gpurun_out/t_parity.log
timeout 600 python bench.py --rows 100000 --steps 2 --warmup 1 --cpu-rows 4096 > gpurun_out/bench_100k.json 2> gpurun_out/bench_100k.err
echo "== bench100k"; tail -5 gpurun_out/bench_100k.err; cat gpurun_out/bench_100k.json
timeout 900 python bench.py --steps 2 --warmup 1 --cpu-rows 8192 > gpurun_out/bench_1m.json 2> gpurun_out/bench_1m.err
echo "== bench1m"; tail -5 gpurun_out/bench_1m.err; cat gpurun_out/bench_1m.json
