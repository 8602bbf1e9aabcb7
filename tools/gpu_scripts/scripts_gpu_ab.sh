#!/bin/bash
# same-box A/B: old library vs new library, alternating, sweep timing only
mkdir -p gpurun_out
for i in 1 2; do
for v in old new; do
  if [ $v = old ]; then export VTF_LIB=$PWD/gpurun_ab/libvtf_old.so; else unset VTF_LIB; fi
  timeout 600 python bench.py --no-cpu-baseline --no-e2e --steps 2 --warmup 1 > gpurun_out/bench_ab_$v$i.json 2> gpurun_out/bench_ab.err; tail -2 gpurun_out/bench_ab.err
  python -c "
import json,sys; r=json.load(open('gpurun_out/bench_ab_$v$i.json')); print('$v', r['ms_per_step'], r['roofline']['achieved'], r['clocks']['sm_mhz'], r['phases_ms']['sil_sweep'])"
done; done
