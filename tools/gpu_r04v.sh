#!/bin/bash
# session r04v: dependent launch (off = sched bit 7) and the warp stagger on the non-TMEM variants (sched bit 6) taken apart at cfg2;
# the earlier A/B (r04e) had both on one bit
mkdir -p gpurun_out
{ for rep in 1 2; do
  echo "== n_fft 1024 batch 32 (cfg2): 3 = PDL on | 131 = PDL off | 67 = PDL on + stagger | 195 = PDL off + stagger"; AB_NFFT=1024 AB_BATCH=32 AB_BURST=20 AB_ROUNDS=7 timeout 300 python tools/ab.py -350,3,0 -350,131,0 -350,67,0 -350,195,0 -200,67,0 -500,67,0
  echo "== n_fft 2048 batch 16 (narrow variant)"; AB_NFFT=2048 AB_BATCH=16 AB_BURST=20 AB_ROUNDS=7 timeout 300 python tools/ab.py -350,3,0 -350,131,0 -350,67,0
  echo "== n_fft 512 batch 64"; AB_NFFT=512 AB_BATCH=64 AB_BURST=20 AB_ROUNDS=7 timeout 300 python tools/ab.py -350,3,0 -350,131,0 -350,67,0
  echo "== n_fft 4096 batch 8 bf16"; AB_NFFT=4096 AB_BATCH=8 AB_DTYPE=bf16 AB_BURST=20 AB_ROUNDS=7 timeout 300 python tools/ab.py -350,3,0 -350,131,0
done; } 2>&1 | tee gpurun_out/r04v_ab_pdl_stagger.txt
