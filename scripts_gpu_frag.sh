#!/bin/bash
# fragment-layout epilogue: layout probe, sweep tests, same-box A/B against the one-row-per-thread epilogue
mkdir -p gpurun_out
timeout 60 tools/probe/tmem_ld_layout > gpurun_out/tmem_ld_layout.txt 2>&1; tail -1 gpurun_out/tmem_ld_layout.txt; sed -n 1,3p gpurun_out/tmem_ld_layout.txt | cut -c1-200
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "sweep or dynamic or silhouette" 2>&1 | tail -5
timeout 600 python tools/sweep_ab.py gpurun_ab/libvtf_frag0.so gpurun_ab/libvtf_frag1.so --sym --modes dynamic --reps 4 2>&1 | tail -2 | cut -c1-420 | tee gpurun_out/sweep_ab_frag.json
