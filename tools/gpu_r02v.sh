#!/bin/bash
mkdir -p gpurun_out
cp fft_b200/_C/libspectre_mix_x0.so fft_b200/_C/libspectre_mix_alt.so
timeout 900 python -m pytest tests -m gpu -q -x --timeout=600 2>&1 | tail -3
{ echo "== smem exchange (SPX_TMEMX=0)"; SPX_ALT=x0 AB_BATCH=148 timeout 300 python tools/ab.py -350,3,0
  echo "== tmem exchange"; AB_BATCH=148 timeout 300 python tools/ab.py -350,3,0 -200,3,0 -100,3,0 0,3,0 -500,3,0 -350,2,0 -350,1,0 -350,0,0 ; } 2>&1 | tee gpurun_out/r02v_ab_tmemx.txt
{ echo "== tmem exchanges"; timeout 300 python tools/timeline.py --batch 64 --skew -350 --sched 3; } > gpurun_out/r02v_timeline.txt 2>&1
grep -A5 "group 0\|group 3" gpurun_out/r02v_timeline.txt | cut -c1-330
