#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
timeout 300 $CMD > gpurun_out/plain.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "== ncu list rc=$?"; wc -l gpurun_out/launches.csv; tail -2 gpurun_out/ncu_list.log | cut -c1-300
