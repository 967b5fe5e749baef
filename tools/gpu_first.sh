#!/bin/bash
# first GPU session: parity tests, smoke, tuning sweep, bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q -x --timeout=600 2>&1 | tail -40 > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log | tail -30
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
timeout 600 python tools/tune.py --out gpurun_out/tune.json > gpurun_out/tune.log 2>&1
tail -45 gpurun_out/tune.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.log 2>&1
tail -3 gpurun_out/bench.log
