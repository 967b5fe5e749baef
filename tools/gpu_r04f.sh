#!/bin/bash
# session r04f: programmatic dependent launch ON by default -- full GPU tests (incl. the dependent-chain test), race hunt, smoke, sanitizer
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x --timeout=600 2>&1 | tail -6 | tee gpurun_out/r04f_pytest_gpu.log
timeout 600 python tools/stress_4096.py 150 11 2>&1 | tail -2 | tee gpurun_out/r04f_stress.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
compute-sanitizer --tool memcheck --error-exitcode 7 python tools/sanitize_small.py > gpurun_out/r04f_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -3 gpurun_out/r04f_memcheck.log
compute-sanitizer --tool synccheck --error-exitcode 7 python tools/sanitize_small.py > gpurun_out/r04f_synccheck.log 2>&1; echo "synccheck rc=$?"; tail -3 gpurun_out/r04f_synccheck.log
