#!/bin/bash
mkdir -p gpurun_out
echo "== tc tests, pair variant (default)"
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "tensor_core or sweep or pick" 2>&1 | tail -30
echo "== tc tests, single-CTA variant"
VTF_SWEEP_VARIANT=1 timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "tensor_core or sweep" 2>&1 | tail -5
echo "== parity"
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu 2>&1 | tail -5
echo "== bench pair"
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_pair.json 2> gpurun_out/bench_pair.err; tail -3 gpurun_out/bench_pair.err; cat gpurun_out/bench_pair.json | python -c "
import json,sys; r=json.loads(sys.stdin.read()); print(r['ms_per_step'], r['value'], r['roofline']['achieved'], r['roofline']['frac'], r['clocks'], r['phases_ms'])"
echo "== bench single"
VTF_SWEEP_VARIANT=1 timeout 600 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/bench_single.json 2> gpurun_out/bench_single.err; tail -3 gpurun_out/bench_single.err; cat gpurun_out/bench_single.json | python -c "
import json,sys; r=json.loads(sys.stdin.read()); print(r['ms_per_step'], r['value'], r['roofline']['achieved'], r['roofline']['frac'], r['clocks'], r['phases_ms'])"
timeout 300 python tools/kernel_bench.py 2>/dev/null | python -c "
import json,sys; r=json.load(sys.stdin)
for k in r['kernels']: print('%-36s %9.3f ms  %8.1f GB/s  frac %.2f  %s' % (k['kernel'], k['ms'], k.get('gbs',0), k.get('hbm_frac',0), ('%.1f TF' % k['tflops']) if 'tflops' in k else ''))
"
