#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout=900 -k "autograd or decode or goldens or metric" 2>&1 | tail -12 > gpurun_out/r02d_pytest_gpu.log
tail -12 gpurun_out/r02d_pytest_gpu.log
timeout 600 python tools/perf_misc.py > gpurun_out/r02d_perf_misc.log 2>&1; cat gpurun_out/r02d_perf_misc.log | tail -20
timeout 300 python tools/sustained.py -350,3,0 -400350,3,0 0,0,0 -350,1,0 -200,3,0 > gpurun_out/r02d_sustained.log 2>&1; cat gpurun_out/r02d_sustained.log
