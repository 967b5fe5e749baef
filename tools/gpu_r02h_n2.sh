#!/bin/bash
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02h_topo.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02h_bench_n2.json 2> gpurun_out/r02h_bench_n2.err
echo "rc=$?"; tail -c 2500 gpurun_out/r02h_bench_n2.json; tail -3 gpurun_out/r02h_bench_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 5 --warmup 2 > gpurun_out/r02h_bench_ref_n2.json 2> gpurun_out/r02h_bench_ref_n2.err
echo "ref rc=$?"; tail -c 600 gpurun_out/r02h_bench_ref_n2.json
