#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -3
for st in 3 2; do
VTF_ROWTILE_STAGES=$st timeout 600 python tools/kernel_bench.py --only "lloyd_assign,classify" > gpurun_out/kb_st$st.json 2>gpurun_out/kb_st$st.err || tail -3 gpurun_out/kb_st$st.err
python - <<PY
import json
r=json.load(open('gpurun_out/kb_st$st.json'))
for x in (r if isinstance(r,list) else r.get('kernels',[])): print('stages=$st', x.get('kernel'), round(x.get('ms',0),4), x.get('tflops'))
PY
done
