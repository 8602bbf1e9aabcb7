#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/kernel_totals.py c4 --scale 0.2 > gpurun_out/totals_c4_200k.json 2> gpurun_out/totals_c4.err; echo rc=$?; tail -3 gpurun_out/totals_c4.err
timeout 600 python tools/kernel_totals.py c3 > gpurun_out/totals_c3_1m.json 2> gpurun_out/totals_c3.err; echo rc=$?; tail -3 gpurun_out/totals_c3.err
