#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x --timeout=900 -k "decode" 2>&1 | tail -4
timeout 300 python tools/perf_misc.py decode 2>&1 | tail -3
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python tools/sanitize_small.py > gpurun_out/r02e_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -6 gpurun_out/r02e_memcheck.log
timeout 900 compute-sanitizer --tool synccheck --error-exitcode 7 python tools/sanitize_small.py > gpurun_out/r02e_synccheck.log 2>&1; echo "synccheck rc=$?"; tail -6 gpurun_out/r02e_synccheck.log
