#!/bin/bash
# bench + ncu launch list of the same command + one full capture of the top kernel
mkdir -p gpurun_out
TAG=${1:-r1}
timeout 900 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
tail -1 gpurun_out/bench_$TAG.json | cut -c1-2500
timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2>> gpurun_out/bench_$TAG.err
tail -1 gpurun_out/bench_ref_$TAG.json | cut -c1-600
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 5 -c 200 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/bench_under_ncu_$TAG.log 2>&1
tail -12 gpurun_out/launches_$TAG.csv | cut -c1-300
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spectre_mix -s 2 -c 1 -f -o gpurun_out/prof_$TAG \
    python tools/prof_one.py --n-fft 4096 --batch 64 2>&1 | tail -3
