#!/bin/bash
mkdir -p gpurun_out
{ for alt in x0 ng0 ng1 ng3 ng0 ng1; do echo "== $alt"; SPX_ALT=$alt timeout 300 python tools/sustained.py -350,3,0; done; } 2>&1 | tee gpurun_out/r02y_sustained_nogate.txt
