#!/bin/bash
# round 2, session r: tensor-memory exchange microbenchmark
mkdir -p gpurun_out
timeout 120 tools/microbench/tmem_xchg > gpurun_out/r02r_tmem_xchg.txt 2>&1
echo "rc=$?" >> gpurun_out/r02r_tmem_xchg.txt
cat gpurun_out/r02r_tmem_xchg.txt
