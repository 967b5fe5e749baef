#!/bin/bash
mkdir -p gpurun_out /tmp/ncu
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spectre_mix -s 2 -c 1 -f -o /tmp/ncu/dit python tools/prof_one.py --n-fft 8192 --batch 32 2>&1 | tail -1
python tools/ncu_summary.py /tmp/ncu/dit.ncu-rep > gpurun_out/r02p_ncu_full_dit2_8192_b32.txt 2>&1
python tools/ncu_lines.py /tmp/ncu/dit.ncu-rep 45 > gpurun_out/r02p_ncu_stall_lines_dit2_8192_b32.txt 2>&1
python tools/ncu_opcodes.py /tmp/ncu/dit.ncu-rep 6144 > gpurun_out/r02p_ncu_opcodes_dit2_8192_b32.txt 2>&1
cat gpurun_out/r02p_ncu_full_dit2_8192_b32.txt | head -36
head -48 gpurun_out/r02p_ncu_stall_lines_dit2_8192_b32.txt
head -20 gpurun_out/r02p_ncu_opcodes_dit2_8192_b32.txt
