#!/bin/bash
# kernel-only rates of the other BASELINE.json configs with the committed library (round-robin bursts, CUDA events)
mkdir -p gpurun_out
{
echo "== cfg2: batch=32 seq=1024 d=768 fp32"
AB_NFFT=1024 AB_BATCH=32 AB_ROUNDS=7 AB_BURST=16 timeout 200 python tools/ab.py -350,3,0
echo "== cfg2 shape at batch=256"
AB_NFFT=1024 AB_BATCH=256 AB_ROUNDS=5 timeout 200 python tools/ab.py -350,3,0
echo "== seq=2048 batch=128"
AB_NFFT=2048 AB_BATCH=128 AB_ROUNDS=5 timeout 200 python tools/ab.py -350,3,0
echo "== metric shape: seq=4096 batch=148 fp32 / bf16"
AB_NFFT=4096 AB_BATCH=148 AB_ROUNDS=5 timeout 200 python tools/ab.py -350,3,0
AB_DTYPE=bf16 AB_NFFT=4096 AB_BATCH=148 AB_ROUNDS=5 timeout 200 python tools/ab.py -350,3,0
echo "== seq=8192 batch=32"
AB_NFFT=8192 AB_BATCH=32 AB_ROUNDS=5 timeout 200 python tools/ab.py -350,3,0
echo "== cfg4: seq=16384 batch=16"
AB_NFFT=16384 AB_BATCH=16 AB_ROUNDS=5 timeout 200 python tools/ab.py -350,3,0
echo "== cfg3: Spectre-base 12 blocks bf16 seq=4096"
timeout 300 python tools/spectre_base_bench.py 8 | tail -1
timeout 300 python tools/spectre_base_bench.py 32 | tail -1
} 2>&1 | tee gpurun_out/configs_r01d.log
