#!/bin/bash
mkdir -p gpurun_out
echo "== dedup tc tests"
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "dedup" 2>&1 | tail -30
echo "== c2"
timeout 600 python tools/config_runs.py c2 2> gpurun_out/cfg_c2.err | tee gpurun_out/cfg_c2.json; tail -3 gpurun_out/cfg_c2.err
echo "== c5 200k dedup"
timeout 600 python tools/config_runs.py c5 --scale 0.125 2> gpurun_out/cfg_c5.err | tee gpurun_out/cfg_c5_500k.json | cut -c1-400; tail -3 gpurun_out/cfg_c5.err
echo "== all tests"
timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -8
