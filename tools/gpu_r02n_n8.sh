#!/bin/bash
mkdir -p gpurun_out
for N in 8 4; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02n_bench_n$N.json 2> gpurun_out/r02n_bench_n$N.err
echo "N=$N rc=$?"; python - <<PY
import json
d=json.loads(open('gpurun_out/r02n_bench_n$N.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms/step',d['ms_per_step'],'frac',d['roofline']['frac'],'burst',d['roofline']['burst']['frac'],'e2e',d['e2e']['value'],'per rank',[round(x/1e6,2) for x in d['e2e']['per_rank_tokens_per_s']],'ceiling',d['e2e']['copy_ceiling']['tokens_per_s'],'clocks',d['clocks'])
PY
tail -2 gpurun_out/r02n_bench_n$N.err
done
