#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/config_dist.py c5 --scale 0.75 > gpurun_out/cfg_c5_3m_1gpu.json 2> gpurun_out/cfg_c5_3m_1gpu.err; echo "c5 rc=$?"; tail -4 gpurun_out/cfg_c5_3m_1gpu.err; cut -c1-1200 gpurun_out/cfg_c5_3m_1gpu.json
CMD="python tools/config_dist.py c4 --scale 0.2"
timeout 600 $CMD > gpurun_out/plain_c4.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:wide_assign -s 3 -c 2 -o gpurun_out/prof_wide_assign $CMD > gpurun_out/ncu_wide.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_wide.log
