#!/bin/bash
# round 2, session t: ncu of the 4096 kernel with the tensor-memory exchanges
mkdir -p gpurun_out /tmp/ncu
cap() {  # tag tiles cmd...
  tag=$1; tiles=$2; shift 2
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:spectre_mix -s 2 -c 1 -f -o /tmp/ncu/$tag "$@" 2>&1 | tail -1
  python tools/ncu_summary.py /tmp/ncu/$tag.ncu-rep > gpurun_out/r02t_ncu_full_$tag.txt 2>&1
  python tools/ncu_lines.py /tmp/ncu/$tag.ncu-rep 60 > gpurun_out/r02t_ncu_stall_lines_$tag.txt 2>&1
  python tools/ncu_opcodes.py /tmp/ncu/$tag.ncu-rep $tiles > gpurun_out/r02t_ncu_opcodes_$tag.txt 2>&1
  head -12 gpurun_out/r02t_ncu_full_$tag.txt
}
cap tmemx_n4096_b64 6144 python tools/prof_one.py --batch 64
