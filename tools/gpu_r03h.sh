#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout=600 2>&1 | tail -3
{ for alt in lz0 "" lz1ng0 lz1ng1 lz0 ""; do echo "== alt='$alt'"; SPX_ALT=$alt AB_BATCH=148 timeout 300 python tools/ab.py -350,3,0 -500,3,0; done; } 2>&1 | tee gpurun_out/r03h_ab_lazy_store_wait.txt
