#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --rows 200000 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
timeout 600 $CMD > gpurun_out/plain.log 2>&1 && timeout 1500 ncu --set full --clock-control none --import-source on -k regex:sil_sweep -s 1 -c 1 -o gpurun_out/prof_sweep_sym $CMD > gpurun_out/ncu_full_sym.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/ncu_full_sym.log | cut -c1-200
