#!/bin/bash
# end-of-round bench lines of the committed state + one full capture of the sub-transform kernel (seq 16384 middle pass)
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/bench_r01d.json 2> gpurun_out/bench_r01d.err
tail -1 gpurun_out/bench_r01d.json | cut -c1-2500
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_r01d.json 2>> gpurun_out/bench_r01d.err
tail -1 gpurun_out/bench_ref_r01d.json | cut -c1-600
timeout 300 ncu --set full --clock-control none --import-source on -k regex:spectre_mix -s 2 -c 1 -f -o gpurun_out/prof_r01d_sub16384 \
    python tools/prof_one.py --n-fft 16384 --batch 16 --reps 4 2>&1 | tail -3
