#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -40 > gpurun_out/t_all.log
echo "== tests"; tail -15 gpurun_out/t_all.log
timeout 600 python tools/kernel_bench.py > gpurun_out/kernels.json 2> gpurun_out/kernels.err
echo "== kernels"; tail -3 gpurun_out/kernels.err; python -c "
import json; r=json.load(open('gpurun_out/kernels.json'))
for k in r['kernels']: print('%-36s %9.3f ms  %8.1f GB/s  frac %.2f  %s' % (k['kernel'], k['ms'], k.get('gbs',0), k.get('hbm_frac',0), ('%.1f TF' % k['tflops']) if 'tflops' in k else ''))
"
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
echo "== bench default"; tail -3 gpurun_out/bench_default.err; cat gpurun_out/bench_default.json
timeout 600 python bench.py --impl reference > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
echo "== bench reference"; tail -3 gpurun_out/bench_reference.err; cat gpurun_out/bench_reference.json | cut -c1-400
CMD="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
timeout 600 $CMD > gpurun_out/plain.log 2>&1 && timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "== ncu list rc=$?"; tail -3 gpurun_out/ncu_list.log; wc -l gpurun_out/launches.csv
CMD2="python bench.py --rows 200000 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
timeout 600 $CMD2 > gpurun_out/plain2.log 2>&1 && timeout 1200 ncu --set full --clock-control none --import-source on -k regex:sil_sweep -c 1 -o gpurun_out/prof_sweep $CMD2 > gpurun_out/ncu_full.log 2>&1
echo "== ncu full rc=$?"; tail -5 gpurun_out/ncu_full.log; ls -la gpurun_out/
