#!/bin/bash
# closing checks of the round on the final tree: full GPU tests, 300-trial race hunt, smoke, the two bench commands the driver runs
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x --timeout=600 2>&1 | tail -3 | tee gpurun_out/r04y_pytest_gpu.log
timeout 900 python tools/stress_4096.py 300 23 2>&1 | tail -2 | tee gpurun_out/r04y_stress.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 | tee gpurun_out/r04y_smoke.log
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r04y_bench_reference_n1.json 2> gpurun_out/r04y_bench_reference_n1.err; tail -c 200 gpurun_out/r04y_bench_reference_n1.json
python bench.py --steps 20 --warmup 5 > gpurun_out/r04y_bench_n1.json 2> gpurun_out/r04y_bench_n1.err; python - <<PY
import json
d=json.loads(open('gpurun_out/r04y_bench_n1.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms/step',d['ms_per_step'],'frac',d['roofline']['frac'],'burst',d['roofline']['burst']['frac'],'parity',d['parity_check']['max_rel_l2'],'e2e',d['e2e']['value'],'ceiling',d['e2e']['copy_ceiling']['tokens_per_s'],'clocks',d['clocks'], 'launch', d['config']['launch'][:40])
PY
