#!/bin/bash
mkdir -p gpurun_out
{ echo "== n_fft 2048 batch 128"; AB_NFFT=2048 AB_BATCH=128 timeout 300 python tools/ab.py -350,3,0 -250,3,0 -450,3,0 -550,3,0 -350,2,0 -350,1,0 0,3,0
  echo "== n_fft 1024 batch 256"; AB_NFFT=1024 AB_BATCH=256 timeout 300 python tools/ab.py -350,3,0 -250,3,0 -450,3,0 -550,3,0 -350,2,0 -350,1,0 0,3,0
  echo "== n_fft 1024 batch 32"; AB_NFFT=1024 AB_BATCH=32 AB_BURST=32 timeout 300 python tools/ab.py -350,3,0 -250,3,0 -150,3,0 0,3,0 -350,2,0
  echo "== n_fft 8192 batch 32"; AB_NFFT=8192 AB_BATCH=32 timeout 300 python tools/ab.py -350,3,0 -250,3,0 -450,3,0 -550,3,0 -350,2,0 -350,1,0 0,3,0; } 2>&1 | tee gpurun_out/r03k_ab_skew_other_sizes.txt
