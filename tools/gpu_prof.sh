#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r1a}
shift
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spectre_mix -s 2 -c 1 -f -o gpurun_out/prof_$TAG python tools/prof_one.py "$@" 2>&1 | tail -5
ls -la gpurun_out/
