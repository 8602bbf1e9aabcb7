#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/gpu_tests.log
timeout 900 python tools/kernel_totals.py c4 > gpurun_out/totals_c4_1m_wide4.json 2> gpurun_out/totals_c4_1m.err; echo "totals rc=$?"; tail -2 gpurun_out/totals_c4_1m.err
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_default.err
CMD="python tools/config_dist.py c4 --scale 0.2"
timeout 600 $CMD > gpurun_out/plain_c4.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:wide_assign -s 3 -c 2 -o gpurun_out/prof_wide_assign2 $CMD > gpurun_out/ncu_wide.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu_wide.log
