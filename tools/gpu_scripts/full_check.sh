#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/gpu_tests.log
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_default.err; python -c "
import json,sys; r=json.load(open('gpurun_out/bench_default.json')); print(r['ms_per_step'], r['value'], 'e2e', r['e2e']['value'], 'roof', r['roofline']['achieved'], r['roofline']['frac'], r['clocks'], r['phases_ms'], r['gpu_launches'])"
