#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -8
timeout 600 python tools/kernel_bench.py > gpurun_out/kernels.json 2>/dev/null; python -c "
import json; r=json.load(open('gpurun_out/kernels.json'))
for k in r['kernels']: print('%-36s %9.3f ms  %8.1f GB/s  frac %.2f  %s' % (k['kernel'], k['ms'], k.get('gbs',0), k.get('hbm_frac',0), ('%.1f TF' % k['tflops']) if 'tflops' in k else ''))
"
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -3 gpurun_out/bench_default.err; python -c "
import json,sys; r=json.load(open('gpurun_out/bench_default.json')); print(r['ms_per_step'], r['value'], 'e2e', r['e2e']['value'], 'roof', r['roofline']['achieved'], r['roofline']['frac'], r['roofline']['traffic'], r['clocks'], r['phases_ms'], r['gpu_launches'])"
CMD="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
timeout 600 $CMD > gpurun_out/plain.log 2>&1 && timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "== ncu list rc=$?"; wc -l gpurun_out/launches.csv
