#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout=600 2>&1 | tail -3
python tools/perf_misc.py shapes 2>&1 | tee gpurun_out/r03j_shapes.log
