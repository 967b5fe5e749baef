#!/bin/bash
# closing checks on the final tree: full GPU suite (incl. the tensor-memory transpose test), race hunt, smoke
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout=900 2>&1 | tail -4 | tee gpurun_out/r03n_pytest_gpu.log
timeout 600 python tools/stress_4096.py 200 7 2>&1 | tail -3 | tee gpurun_out/r03n_stress.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
