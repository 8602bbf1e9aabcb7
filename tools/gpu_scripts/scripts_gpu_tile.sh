#!/bin/bash
# register-tiled assign / classify: kernel + parity tests, per-kernel timings with and without the fast path
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -5
timeout 600 python tools/kernel_bench.py --only "lloyd_assign,classify,cluster_sums,dispersion" > gpurun_out/kb_tile.json 2> gpurun_out/kb_tile.err; tail -2 gpurun_out/kb_tile.err
python - <<'PY'
import json
r=json.load(open('gpurun_out/kb_tile.json'))
for x in (r if isinstance(r,list) else r.get('kernels',[])): print(x.get('kernel'), round(x.get('ms',0),4), x.get('tflops'))
PY
VTF_ROWTILE=0 timeout 600 python tools/kernel_bench.py --only "lloyd_assign,classify" > gpurun_out/kb_notile.json 2>/dev/null
python - <<'PY'
import json
r=json.load(open('gpurun_out/kb_notile.json'))
for x in (r if isinstance(r,list) else r.get('kernels',[])): print('rowdot only:', x.get('kernel'), round(x.get('ms',0),4), x.get('tflops'))
PY
