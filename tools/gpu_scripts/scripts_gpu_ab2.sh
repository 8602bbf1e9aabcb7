#!/bin/bash
echo "== SYM"; python tools/sweep_ab.py --sym --reps 3 gpurun_ab/libvtf_e8.so gpurun_ab/libvtf_e16.so gpurun_ab/libvtf_e16r.so gpurun_ab/libvtf_e8r.so 2>&1 | tail -5 | cut -c1-200
echo "== FULL"; python tools/sweep_ab.py --reps 2 gpurun_ab/libvtf_e8.so gpurun_ab/libvtf_e16.so 2>&1 | tail -3 | cut -c1-200
