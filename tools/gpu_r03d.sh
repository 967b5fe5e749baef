#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/timeline.py --n-fft 8192 --batch 32 --skew -350 --sched 3 > gpurun_out/r03d_timeline_8192.txt 2>&1
grep -A5 "group 0\|group 3" gpurun_out/r03d_timeline_8192.txt | cut -c1-330
