#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/config_runs.py c2 > gpurun_out/cfg_c2.json 2> gpurun_out/cfg_c2.err; tail -3 gpurun_out/cfg_c2.err; cat gpurun_out/cfg_c2.json
timeout 900 python tools/config_runs.py c4 --scale 0.2 > gpurun_out/cfg_c4_200k.json 2> gpurun_out/cfg_c4.err; tail -3 gpurun_out/cfg_c4.err; cat gpurun_out/cfg_c4_200k.json
timeout 900 python tools/config_runs.py c5 --scale 0.125 > gpurun_out/cfg_c5_500k.json 2> gpurun_out/cfg_c5.err; tail -3 gpurun_out/cfg_c5.err; cat gpurun_out/cfg_c5_500k.json
