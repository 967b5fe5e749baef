#!/bin/bash
# compute warps alone (sched bit 3: no tile I/O, results invalid) and tile I/O alone (bit 2) for the no-gate builds
mkdir -p gpurun_out
{ for alt in "" ng0 ng1 ng3; do echo "== alt='$alt'  (sched 3 = full kernel, 11 = FFT passes only, 7 = tile I/O only)"; SPX_ALT=$alt AB_BATCH=148 timeout 300 python tools/ab.py -350,3,0 -350,11,0 -350,7,0; done; } 2>&1 | tee gpurun_out/r03l_ab_compute_only.txt
