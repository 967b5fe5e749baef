#!/bin/bash
# A/B of the sub-transform variant's stage-0 twiddle source (shared memory vs L2) on the long-context path, then the GPU tests
mkdir -p gpurun_out
{
echo "== 16384 B=16 stage-0 twiddles in shared memory (default build)"
AB_NFFT=16384 AB_BATCH=16 AB_ROUNDS=7 timeout 200 python tools/ab.py -350,3,0
echo "== 16384 B=16 stage-0 twiddles through L2 (SPX_SUB_TW_SMEM=0 build)"
SPX_ALT=subl2 AB_NFFT=16384 AB_BATCH=16 AB_ROUNDS=7 timeout 200 python tools/ab.py -350,3,0
echo "== again, default"
AB_NFFT=16384 AB_BATCH=16 AB_ROUNDS=7 timeout 200 python tools/ab.py -350,3,0
echo "== 8192 B=32 default / alt"
AB_NFFT=8192 AB_BATCH=32 AB_ROUNDS=5 timeout 200 python tools/ab.py -350,3,0
SPX_ALT=subl2 AB_NFFT=8192 AB_BATCH=32 AB_ROUNDS=5 timeout 200 python tools/ab.py -350,3,0
} 2>&1 | tee gpurun_out/ab_sub.log
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.log
