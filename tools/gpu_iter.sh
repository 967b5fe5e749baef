#!/bin/bash
# iteration session: parity tests (stop at first failure), tuning sweep, short bench
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x --timeout=600 2>&1 | tail -25 > gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
timeout 600 python tools/tune.py --out gpurun_out/tune.json > gpurun_out/tune.log 2>&1
grep -v '^$' gpurun_out/tune.log | tail -60
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/bench.log 2>&1
tail -2 gpurun_out/bench.log | cut -c1-1200
