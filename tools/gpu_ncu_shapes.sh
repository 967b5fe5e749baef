#!/bin/bash
# ncu time / DRAM bytes / pipe utilisation of the kernels behind the other BASELINE shapes (seq 1024, seq 16384 two-pass)
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__cycles_active.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread,launch__block_size,launch__grid_size
for cfg in "1024 256" "16384 16"; do
  set -- $cfg
  timeout 300 ncu --metrics $M --clock-control none -k regex:spectre\|long_pass -s 6 -c 3 --csv --log-file gpurun_out/shape_$1.csv python tools/prof_one.py --n-fft $1 --batch $2 --reps 4 > /dev/null 2>&1
  echo "== n_fft $1 batch $2 (C=768, d_g=16, fp32)"; python - <<PY
import csv
rows = [r for r in csv.reader(open("gpurun_out/shape_$1.csv")) if len(r) > 10]
cur = None
for r in rows[1:]:
    if r[0] != cur:
        cur = r[0]; print("kernel", r[4][:110])
    print(f"   {r[-3]:75s} {r[-1]}")
PY
done
