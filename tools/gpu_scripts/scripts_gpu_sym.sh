#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "symmetric or tensor_core_sweep" 2>&1 | tail -25
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_sym.json 2> gpurun_out/bench_sym.err; tail -3 gpurun_out/bench_sym.err; python -c "
import json,sys; r=json.load(open('gpurun_out/bench_sym.json')); print(r['ms_per_step'], r['value'], 'e2e', r['e2e']['value'], 'roof', r['roofline']['achieved'], r['roofline']['frac'], r['clocks'], r['phases_ms'], r['scores'])"
