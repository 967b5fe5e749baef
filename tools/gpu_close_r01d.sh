#!/bin/bash
# closing check of the round: GPU tests, smoke(), ncu launch list of the bench command as committed
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 5 -c 200 --csv --log-file gpurun_out/launches_r01d.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/bench_under_ncu_r01d.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/launches_r01d.csv")) if len(r) > 10 and r[0].isdigit()]
t = collections.defaultdict(float); n = collections.Counter()
for r in rows:
    k = r[4][:90]; t[k] += float(r[-1]); n[k] += 1
tot = sum(t.values())
for k, v in sorted(t.items(), key=lambda x: -x[1])[:8]:
    print(f"{100 * v / tot:5.1f}%  {n[k]:4d}x  {k}")
PY
