#!/bin/bash
# round 2, first GPU session: parity tests, DSMEM microbenchmark, bench line (cfg5 workload) and the reference arm
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/r02a_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -x --timeout=900 2>&1 | tail -30 > gpurun_out/r02a_pytest_gpu.log
tail -30 gpurun_out/r02a_pytest_gpu.log
timeout 120 tools/microbench/dsmem_rates > gpurun_out/r02a_dsmem.log 2>&1
cat gpurun_out/r02a_dsmem.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err
echo "bench rc=$?"; tail -c 3000 gpurun_out/r02a_bench.json; tail -5 gpurun_out/r02a_bench.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/r02a_bench_ref.json 2> gpurun_out/r02a_bench_ref.err
tail -c 1500 gpurun_out/r02a_bench_ref.json
