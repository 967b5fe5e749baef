#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout=600 2>&1 | tail -3
bash tools/gpu_r02o.sh 2>&1 | tee gpurun_out/r03e_dit2_async_gate.txt
timeout 300 python tools/timeline.py --n-fft 8192 --batch 32 --skew -350 --sched 3 > gpurun_out/r03e_timeline_8192.txt 2>&1
grep "tile#4" gpurun_out/r03e_timeline_8192.txt | cut -c1-330
