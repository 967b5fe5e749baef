#!/bin/bash
# round 2 ncu evidence: launch list of the bench command, full captures of the metric kernel, the gate-gradient kernel and the
# n_fft = 8192 single kernel (summaries extracted on the box: the reports themselves are 20-30 MB each); SHFL-vs-LDS microbenchmark
mkdir -p gpurun_out /tmp/ncu
tools/microbench/shfl_vs_lds > gpurun_out/r02i_shfl_vs_lds.txt 2>&1; cat gpurun_out/r02i_shfl_vs_lds.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 480 --csv --log-file gpurun_out/r02i_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/r02i_bench_under_ncu.log 2>&1
wc -l gpurun_out/r02i_launches_bench.csv
cap() {  # tag tiles cmd...
  tag=$1; tiles=$2; shift 2
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:spectre_mix -s 2 -c 1 -f -o /tmp/ncu/$tag "$@" 2>&1 | tail -1
  python tools/ncu_summary.py /tmp/ncu/$tag.ncu-rep > gpurun_out/r02i_ncu_full_$tag.txt 2>&1
  python tools/ncu_lines.py /tmp/ncu/$tag.ncu-rep 40 > gpurun_out/r02i_ncu_stall_lines_$tag.txt 2>&1
  python tools/ncu_opcodes.py /tmp/ncu/$tag.ncu-rep $tiles > gpurun_out/r02i_ncu_opcodes_$tag.txt 2>&1
  head -12 gpurun_out/r02i_ncu_full_$tag.txt
}
cap n4096_b64 6144 python tools/prof_one.py --batch 64
cap n8192_b32 6144 python tools/prof_one.py --n-fft 8192 --batch 32
cat > /tmp/dg.py <<'PY'
import torch, sys, os
sys.path.insert(0, os.getcwd())
from fft_b200 import ops
V = torch.randn(64, 4096, 768, device='cuda'); dY = torch.randn(64, 4096, 768, device='cuda')
for _ in range(4): g = ops._dgate_fused(V, dY, 4096, 16)
torch.cuda.synchronize(); print('ok', float(g.real.abs().mean()))
PY
cap dgate_b64 6144 python /tmp/dg.py
cp /tmp/ncu/n4096_b64.ncu-rep gpurun_out/prof_r02i_n4096_b64.ncu-rep
du -sh gpurun_out
