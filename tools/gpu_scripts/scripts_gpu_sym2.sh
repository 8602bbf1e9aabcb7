#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "symmetric or tensor_core_sweep" 2>&1 | tail -3
for i in 1 2; do timeout 600 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/bench_q$i.json 2> gpurun_out/bench_q.err; tail -2 gpurun_out/bench_q.err; python -c "
import json,sys; r=json.load(open('gpurun_out/bench_q$i.json')); print(r['ms_per_step'], r['value'], r['roofline']['achieved'], r['roofline']['frac'], r['clocks'], r['phases_ms']['sil_sweep'])"; done
