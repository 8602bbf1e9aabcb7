#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "pick" 2>&1 | tail -30
timeout 300 python tools/kernel_bench.py --only kpp_pick 2>/dev/null | python -c "
import json,sys; r=json.load(sys.stdin)
for k in r['kernels']: print('%-36s %9.3f ms' % (k['kernel'], k['ms']))"
