#!/bin/bash
# Decomposition of the n_fft = 4096 kernel (GPU box): full kernel / FFT passes only (sched bit 3) / tile I/O only (sched bit 2),
# round-robin timed (tools/ab.py).  The two diagnostic modes produce invalid results by design.
mkdir -p gpurun_out
python tools/ab.py -350,3,0 -350,11,0 -350,7,0 0,0,0 2>&1 | tee gpurun_out/diag.log
