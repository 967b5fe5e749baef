#!/bin/bash
mkdir -p gpurun_out
cp fft_b200/_C/libspectre_mix_x0.so fft_b200/_C/libspectre_mix_alt.so
{ echo "== smem exchanges"; SPX_ALT=1 timeout 300 python tools/timeline.py --batch 64 --skew -350 --sched 3
  echo "== tmem exchanges"; timeout 300 python tools/timeline.py --batch 64 --skew -350 --sched 3; } > gpurun_out/r02u_timeline.txt 2>&1
cat gpurun_out/r02u_timeline.txt
