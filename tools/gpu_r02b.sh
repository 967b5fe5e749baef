#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout=900 2>&1 | tail -30 > gpurun_out/r02b_pytest_gpu.log
tail -30 gpurun_out/r02b_pytest_gpu.log
timeout 300 python tools/ab_anchors.py > gpurun_out/r02b_ab_anchors.log 2>&1; tail -3 gpurun_out/r02b_ab_anchors.log
AB_BATCH=32 AB_NFFT=1024 timeout 300 python tools/ab_anchors.py >> gpurun_out/r02b_ab_anchors.log 2>&1; tail -1 gpurun_out/r02b_ab_anchors.log
