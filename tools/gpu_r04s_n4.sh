#!/bin/bash
# final tree on 4 GPUs: reference arm, then our arm, as the driver launches them
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 4 --steps 5 --warmup 2 > gpurun_out/r04s_bench_reference_n4.json 2> gpurun_out/r04s_bench_reference_n4.err
echo "ref rc=$?"; tail -c 300 gpurun_out/r04s_bench_reference_n4.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r04s_bench_n4.json 2> gpurun_out/r04s_bench_n4.err
echo "rc=$?"; python - <<PY
import json
d=json.loads(open('gpurun_out/r04s_bench_n4.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms/step',d['ms_per_step'],'frac',d['roofline']['frac'],'burst',d['roofline']['burst']['frac'],'parity',d['parity_check']['max_rel_l2'],'e2e',d['e2e']['value'],'per rank',[round(x/1e6,2) for x in d['e2e']['per_rank_tokens_per_s']],'ceiling',d['e2e']['copy_ceiling']['tokens_per_s'],'clocks',d['clocks'])
PY
tail -2 gpurun_out/r04s_bench_n4.err
