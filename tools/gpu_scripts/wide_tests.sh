#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_wide.py -x -q > gpurun_out/wide_tests.log 2>&1; echo "rc=$?"; tail -40 gpurun_out/wide_tests.log
