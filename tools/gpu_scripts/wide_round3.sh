#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_wide.py tests/test_gpu_parity.py -x -q > gpurun_out/wide_tests.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/wide_tests.log
timeout 900 python tools/kernel_totals.py c4 > gpurun_out/totals_c4_1m_wide3.json 2> gpurun_out/totals_c4_1m.err; echo rc=$?; tail -3 gpurun_out/totals_c4_1m.err
