#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
timeout 600 $CMD > gpurun_out/plain.log 2>&1 && timeout 1500 ncu --set full --clock-control none --import-source on -k regex:sil_sweep -s 1 -c 1 -o gpurun_out/prof_sweep_1m $CMD > gpurun_out/ncu_full_1m.log 2>&1
echo "full rc=$?"
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "list rc=$?"; wc -l gpurun_out/launches.csv
