#!/bin/bash
# ncu pipe utilisation / stall ratios of the 4096 kernel in its three modes: full, FFT passes only, tile I/O only
mkdir -p gpurun_out
M=gpu__time_duration.sum,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed.sum,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio,smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio,dram__cycles_active.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed
for sched in 3 11 7; do
  timeout 300 ncu --metrics $M --clock-control none -k regex:spectre_mix -s 2 -c 1 --csv --log-file gpurun_out/modes_$sched.csv python tools/prof_one.py --n-fft 4096 --batch 64 --sched $sched > /dev/null 2>&1
  echo "== sched $sched"; python - <<PY
import csv
rows = [r for r in csv.reader(open("gpurun_out/modes_$sched.csv")) if len(r) > 10]
for r in rows[1:]:
    print(f"{r[-3]:90s} {r[-1]}")
PY
done
