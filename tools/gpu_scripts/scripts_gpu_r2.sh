#!/bin/bash
# sync-free k-means++: kernel tests + parity tests, then 2-rank dist_check and bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x -k "kpp or cluster or kmeans or lloyd or seeds" 2>&1 | tail -5
bash scripts_gpu_multi.sh 2 2>&1 | cut -c1-2500
