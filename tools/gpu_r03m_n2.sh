#!/bin/bash
# closing run on 2 GPUs: reference arm, then our arm, as the driver launches them
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 5 --warmup 2 > gpurun_out/r03m_bench_reference_n2.json 2> gpurun_out/r03m_bench_reference_n2.err
echo "ref rc=$?"; tail -c 400 gpurun_out/r03m_bench_reference_n2.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r03m_bench_n2.json 2> gpurun_out/r03m_bench_n2.err
echo "rc=$?"; tail -c 3000 gpurun_out/r03m_bench_n2.json; tail -3 gpurun_out/r03m_bench_n2.err
