#!/bin/bash
# final-tree ncu captures of the kernels beside the metric kernel: decode step (both kernels), n_fft 8192 (DIT2), 1024 x batch 32 (cfg2), bf16 4096
mkdir -p gpurun_out /tmp/ncu
cap() {  # tag kernel-regex skip cmd...
  tag=$1; rx=$2; skip=$3; shift 3
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -f -o /tmp/ncu/$tag "$@" 2>&1 | tail -1
  python tools/ncu_summary.py /tmp/ncu/$tag.ncu-rep > gpurun_out/r04u_ncu_full_$tag.txt 2>&1
  head -8 gpurun_out/r04u_ncu_full_$tag.txt
}
cap decode_step decode_kernel 2 python tools/decode_one.py
cap decode_reduce decode_reduce 2 python tools/decode_one.py
cap n8192_b32 spectre_mix 2 python tools/prof_one.py --n-fft 8192 --batch 32
cap n1024_b32 spectre_mix 2 python tools/prof_one.py --n-fft 1024 --batch 32
cap n4096_b148_bf16 spectre_mix 2 python tools/prof_one.py --batch 148 --bf16
