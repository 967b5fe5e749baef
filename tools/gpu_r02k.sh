#!/bin/bash
mkdir -p gpurun_out
for B in 16 32 48 64 96 128 192; do for tile in 0 16; do
  AB_ROUNDS=3 AB_NFFT=1024 AB_BATCH=$B AB_TILE=$tile timeout 200 python tools/ab.py -350,3,0 2>&1 | tail -1 | sed "s/^/n_fft 1024 B $B tile_override $tile: /" | cut -c1-150
done; done | tee gpurun_out/r02k_ab_wide_1024.log
for B in 8 16 24 32 48 64 96; do for tile in 0 8; do
  AB_ROUNDS=3 AB_NFFT=2048 AB_BATCH=$B AB_TILE=$tile timeout 200 python tools/ab.py -350,3,0 2>&1 | tail -1 | sed "s/^/n_fft 2048 B $B tile_override $tile: /" | cut -c1-150
done; done | tee gpurun_out/r02k_ab_wide_2048.log
