#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -5
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -3 gpurun_out/bench_default.err; python -c "
import json,sys; r=json.load(open('gpurun_out/bench_default.json')); print(r['ms_per_step'], r['value'], 'e2e', r['e2e']['value'], 'roof', r['roofline']['achieved'], r['roofline']['frac'], r['clocks'], r['phases_ms'], r['gpu_launches'], r['scores'])"
