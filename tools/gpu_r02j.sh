#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout=900 2>&1 | tail -12 > gpurun_out/r02j_pytest_gpu.log
tail -12 gpurun_out/r02j_pytest_gpu.log
timeout 300 python tools/perf_misc.py shapes > gpurun_out/r02j_shapes.log 2>&1; cat gpurun_out/r02j_shapes.log
for cfg in "1024 32" "1024 256" "2048 128"; do set -- $cfg
  for tile in 0 16; do
    AB_NFFT=$1 AB_BATCH=$2 AB_TILE=$tile timeout 200 python tools/ab.py -350,3,0 2>&1 | tail -1 | sed "s/^/n_fft $1 B $2 tile_override $tile: /"
  done
done | tee gpurun_out/r02j_ab_wide.log
