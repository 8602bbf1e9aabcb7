#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_wide.py -x -q > gpurun_out/wide_tests.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/wide_tests.log
VTF_WIDE_MIN_SUM_K=8 timeout 600 python tools/kernel_totals.py c3 > gpurun_out/totals_c3_1m_wide.json 2> gpurun_out/totals_c3.err; echo rc=$?; tail -2 gpurun_out/totals_c3.err
