#!/bin/bash
mkdir -p gpurun_out
SPX_ALT=1 timeout 300 python tools/timeline.py --batch 64 --skew -350 --sched 3 > gpurun_out/r03g_timeline_helper.txt 2>&1
grep -A6 "helper warpgroup" gpurun_out/r03g_timeline_helper.txt | cut -c1-330; grep "tile#4" gpurun_out/r03g_timeline_helper.txt | cut -c1-300
