#!/bin/bash
# chunked (L2-resident column chunk) symmetric sweep, static list vs work counter, chunk sizes; then bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "sweep or dynamic or silhouette" 2>&1 | tail -5
LIB=video-to-faces_b200/videotofaces_b200/_lib/libvtf_b200.so
for c in 64 32 128; do
  echo "== chunk $c"
  VTF_SWEEP_CHUNK=$c timeout 600 python tools/sweep_ab.py $LIB --sym --modes static,dynamic --reps 3 2>&1 | tail -2 | cut -c1-400 | tee -a gpurun_out/sweep_ab_dyn.json
done
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -3 gpurun_out/bench_default.err; python -c "
import json,sys; r=json.load(open('gpurun_out/bench_default.json')); print(r['ms_per_step'], r['value'], 'e2e', r['e2e']['value'], 'roof', r['roofline']['achieved'], r['roofline']['frac'], r['clocks'], r['phases_ms'], r['gpu_launches'])"
