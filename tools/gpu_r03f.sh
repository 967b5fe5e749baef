#!/bin/bash
# final tree of the round: sanitizer on every kernel family, shapes table, the bench lines the driver runs
mkdir -p gpurun_out
compute-sanitizer --tool memcheck --error-exitcode 7 python tools/sanitize_small.py > gpurun_out/r03f_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -3 gpurun_out/r03f_memcheck.log
compute-sanitizer --tool synccheck --error-exitcode 7 python tools/sanitize_small.py > gpurun_out/r03f_synccheck.log 2>&1; echo "synccheck rc=$?"; tail -3 gpurun_out/r03f_synccheck.log
python tools/perf_misc.py shapes > gpurun_out/r03f_shapes.log 2>&1; cat gpurun_out/r03f_shapes.log
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r03f_bench_reference_n1.json 2> gpurun_out/r03f_bench_reference_n1.err; tail -c 600 gpurun_out/r03f_bench_reference_n1.json
python bench.py --steps 20 --warmup 5 > gpurun_out/r03f_bench_n1.json 2> gpurun_out/r03f_bench_n1.err; cat gpurun_out/r03f_bench_n1.json
