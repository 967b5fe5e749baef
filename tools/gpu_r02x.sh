#!/bin/bash
mkdir -p gpurun_out
{ for alt in x0 ng0 ng1 ng3; do echo "== $alt"; SPX_ALT=$alt AB_BATCH=148 timeout 300 python tools/ab.py -350,3,0 -200,3,0 -500,3,0; done; } 2>&1 | tee gpurun_out/r02x_ab_nogate.txt
