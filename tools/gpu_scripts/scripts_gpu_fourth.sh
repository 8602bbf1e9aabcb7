#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -40 > gpurun_out/t_all.log
echo "== tests"; tail -25 gpurun_out/t_all.log
timeout 300 python tools/kernel_bench.py 2>/dev/null > gpurun_out/kernels.json; python -c "
import json; r=json.load(open('gpurun_out/kernels.json'))
for k in r['kernels']: print('%-36s %9.3f ms  %8.1f GB/s  frac %.2f  %s' % (k['kernel'], k['ms'], k.get('gbs',0), k.get('hbm_frac',0), ('%.1f TF' % k['tflops']) if 'tflops' in k else ''))
"
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_v3.json 2> gpurun_out/bench_v3.err; tail -3 gpurun_out/bench_v3.err; cat gpurun_out/bench_v3.json | python -c "
import json,sys; r=json.loads(sys.stdin.read()); print(r['ms_per_step'], r['value'], r['e2e'], r['roofline']['achieved'], r['roofline']['frac'], r['clocks'], r['phases_ms'], r['gpu_launches'])"
