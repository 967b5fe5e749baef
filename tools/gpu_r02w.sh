#!/bin/bash
mkdir -p gpurun_out
{ echo "== tmem exchanges, (diagnostic marks of that session: built with a since-removed -DSPX_DIAG_X)"; SPX_ALT=1 timeout 300 python tools/timeline.py --batch 64 --skew -350 --sched 3; } > gpurun_out/r02w_timeline.txt 2>&1
grep -A5 "group 0\|group 1\|group 2\|group 3" gpurun_out/r02w_timeline.txt | cut -c1-330
