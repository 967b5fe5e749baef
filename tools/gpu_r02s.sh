#!/bin/bash
# round 2, session s: stage-1 <-> middle-pass exchanges through tensor memory -- parity, then A/B against the shared-memory build
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout=600 2>&1 | tail -5 > gpurun_out/r02s_pytest_gpu.log
cat gpurun_out/r02s_pytest_gpu.log
for rep in 1 2; do
  echo "== smem exchange (SPX_TMEMX=0)"; SPX_ALT=x0 AB_BATCH=148 timeout 300 python tools/ab.py -350,3,0 -200,3,0 0,3,0
  echo "== tmem exchange"; AB_BATCH=148 timeout 300 python tools/ab.py -350,3,0 -200,3,0 -100,3,0 0,3,0 -500,3,0 -350,2,0 -350,0,0
done 2>&1 | tee gpurun_out/r02s_ab_tmemx.txt
