#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout=900 2>&1 | tail -5 > gpurun_out/r02q_pytest_gpu.log
tail -5 gpurun_out/r02q_pytest_gpu.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python tools/sanitize_small.py > gpurun_out/r02q_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/r02q_memcheck.log
timeout 900 compute-sanitizer --tool synccheck --error-exitcode 7 python tools/sanitize_small.py > gpurun_out/r02q_synccheck.log 2>&1; echo "synccheck rc=$?"; tail -4 gpurun_out/r02q_synccheck.log
timeout 300 python tools/perf_misc.py shapes > gpurun_out/r02q_shapes.log 2>&1; cat gpurun_out/r02q_shapes.log
