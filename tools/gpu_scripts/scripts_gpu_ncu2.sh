#!/bin/bash
mkdir -p gpurun_out
CMD="python tools/kernel_bench.py --only lloyd_assign,kpp_dist,cluster_sums,kpp_pick,classify --reps 1 --warm 0"
timeout 300 $CMD > gpurun_out/kb_plain.json 2> gpurun_out/kb_plain.err && timeout 900 ncu --set full --clock-control none --import-source on -k regex:"rowdot_kernel|kpp_dist_kernel|cluster_sums_kernel|kpp_pick_kernel" -c 8 -o gpurun_out/prof_small $CMD > gpurun_out/ncu_small.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/ncu_small.log; ls -la gpurun_out/prof_small.ncu-rep
