#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/kernel_totals.py c4 --scale 0.2 > gpurun_out/totals_c4_200k_wide.json 2> gpurun_out/totals_c4.err; echo rc=$?; tail -3 gpurun_out/totals_c4.err
timeout 900 python tools/kernel_totals.py c4 > gpurun_out/totals_c4_1m_wide.json 2> gpurun_out/totals_c4_1m.err; echo rc=$?; tail -3 gpurun_out/totals_c4_1m.err
