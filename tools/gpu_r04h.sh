#!/bin/bash
# session r04h: final tree of the round -- the two bench lines the driver runs, the shapes table, launch list + full capture of the metric kernel
mkdir -p gpurun_out /tmp/ncu
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r04h_bench_reference_n1.json 2> gpurun_out/r04h_bench_reference_n1.err; tail -c 400 gpurun_out/r04h_bench_reference_n1.json
python bench.py --steps 20 --warmup 5 > gpurun_out/r04h_bench_n1.json 2> gpurun_out/r04h_bench_n1.err; cat gpurun_out/r04h_bench_n1.json
python tools/perf_misc.py shapes > gpurun_out/r04h_shapes_kernel_only.txt 2>&1; cat gpurun_out/r04h_shapes_kernel_only.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 480 --csv --log-file gpurun_out/r04h_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/r04h_bench_under_ncu.log 2>&1
wc -l gpurun_out/r04h_launches_bench.csv
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spectre_mix -s 2 -c 1 -f -o /tmp/ncu/n4096_b64 python tools/prof_one.py --batch 64 2>&1 | tail -1
python tools/ncu_summary.py /tmp/ncu/n4096_b64.ncu-rep > gpurun_out/r04h_ncu_full_n4096_b64.txt 2>&1
head -14 gpurun_out/r04h_ncu_full_n4096_b64.txt
