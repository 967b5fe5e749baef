#!/bin/bash
# session r04e: programmatic dependent launch (sched bit 6) -- full GPU tests with the knob on, then A/B at cfg2 and the metric shape
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout=600 2>&1 | tail -4 | tee gpurun_out/r04e_pytest_gpu.log
{ for rep in 1 2; do
  echo "== n_fft 1024 batch 32 (cfg2), tile override 16"; AB_NFFT=1024 AB_BATCH=32 AB_TILE=16 AB_BURST=20 AB_ROUNDS=7 timeout 300 python tools/ab.py -350,3,0 -350,67,0
  echo "== n_fft 1024 batch 32 (cfg2), default variant"; AB_NFFT=1024 AB_BATCH=32 AB_BURST=20 AB_ROUNDS=7 timeout 300 python tools/ab.py -350,3,0 -350,67,0
  echo "== n_fft 4096 batch 8 bf16 (cfg3's launch)"; AB_NFFT=4096 AB_BATCH=8 AB_DTYPE=bf16 AB_BURST=20 AB_ROUNDS=7 timeout 300 python tools/ab.py -350,3,0 -350,67,0
  echo "== n_fft 4096 batch 148"; AB_BATCH=148 timeout 300 python tools/ab.py -350,3,0 -350,67,0
  echo "== n_fft 2048 batch 16"; AB_NFFT=2048 AB_BATCH=16 AB_BURST=20 AB_ROUNDS=7 timeout 300 python tools/ab.py -350,3,0 -350,67,0
  echo "== n_fft 8192 batch 8"; AB_NFFT=8192 AB_BATCH=8 AB_BURST=20 AB_ROUNDS=7 timeout 300 python tools/ab.py -350,3,0 -350,67,0
done; } 2>&1 | tee gpurun_out/r04e_ab_pdl.txt
