#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout=900 2>&1 | tail -8 > gpurun_out/r02g_pytest_gpu.log
tail -8 gpurun_out/r02g_pytest_gpu.log
timeout 300 python tools/perf_misc.py shapes > gpurun_out/r02g_shapes.log 2>&1; cat gpurun_out/r02g_shapes.log
