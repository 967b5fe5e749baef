#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout=900 2>&1 | tail -6 > gpurun_out/r02m_pytest_gpu.log
tail -6 gpurun_out/r02m_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
AB_NFFT=16384 AB_BATCH=16 timeout 300 python tools/ab.py -350,3,0 2>&1 | tail -1
