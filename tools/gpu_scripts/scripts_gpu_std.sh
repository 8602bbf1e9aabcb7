#!/bin/bash
# standard round-end style check: all GPU tests, smoke, default bench, reference arm
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -6
timeout 600 python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -3 gpurun_out/bench_default.err; python -c "
import json,sys; r=json.load(open('gpurun_out/bench_default.json')); print(r['ms_per_step'], r['value'], 'e2e', r['e2e']['value'], 'roof', r['roofline']['achieved'], r['roofline']['frac'], r['clocks'], r['phases_ms'], 'cpu', r['cpu_baseline']['value'], r['gpu_launches'])"
