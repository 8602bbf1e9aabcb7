#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/gpu_tests.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log | cut -c1-300
timeout 600 python tools/kernel_totals.py c3 > gpurun_out/totals_c3_1m.json 2> gpurun_out/totals_c3.err; echo rc=$?
timeout 900 python tools/kernel_totals.py c4 > gpurun_out/totals_c4_1m_wide5.json 2> gpurun_out/totals_c4_1m.err; echo rc=$?
