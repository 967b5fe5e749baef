#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout=900 2>&1 | tail -4 > gpurun_out/r02l_pytest_gpu.log
tail -4 gpurun_out/r02l_pytest_gpu.log
python -c "
import fft_b200
for B in (16,32,48,64,256): print(1024, B, fft_b200.plan_info(B,1024,1024,768,16)['tile_channels'])
for B in (8,16,24,32,128): print(2048, B, fft_b200.plan_info(B,2048,2048,768,16)['tile_channels'])
"
timeout 300 python tools/perf_misc.py shapes > gpurun_out/r02l_shapes.log 2>&1; cat gpurun_out/r02l_shapes.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02l_bench_n1.json 2> gpurun_out/r02l_bench_n1.err
echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02l_bench_n1.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms/step',d['ms_per_step'],'frac',d['roofline']['frac'],'burst',d['roofline']['burst']['frac'],'e2e',d['e2e']['value'],'ceiling',d['e2e']['copy_ceiling']['tokens_per_s'],'cpu',d['cpu_baseline']['value'],'clocks',d['clocks'],'parity',d['parity_check']['max_rel_l2'])
PY
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02l_bench_ref_n1.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/r02l_bench_ref_n1.json').read().strip().splitlines()[-1]); print('reference', d['value'], d['cpu_baseline']['stock_reference_module'])"
timeout 300 python tools/spectre_base_bench.py 2>&1 | tail -6
