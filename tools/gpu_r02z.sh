#!/bin/bash
mkdir -p gpurun_out
{ for alt in pf0 pf1 pf0 pf1; do echo "== $alt"; SPX_ALT=$alt AB_BATCH=148 timeout 300 python tools/ab.py -350,3,0 -500,3,0; done
  for alt in pf0 pf1; do echo "== sustained $alt"; SPX_ALT=$alt timeout 300 python tools/sustained.py -350,3,0; done; } 2>&1 | tee gpurun_out/r02z_ab_gate_prefetch.txt
