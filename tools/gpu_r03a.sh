#!/bin/bash
# helper-staged gate rows: parity, A/B against the compute-staged build, sustained
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout=600 2>&1 | tail -3
{ for alt in hg0 "" hg0 ""; do echo "== alt='$alt'"; SPX_ALT=$alt AB_BATCH=148 timeout 300 python tools/ab.py -350,3,0 -500,3,0 -200,3,0; done
  for alt in hg0 ""; do echo "== sustained alt='$alt'"; SPX_ALT=$alt timeout 300 python tools/sustained.py -350,3,0; done; } 2>&1 | tee gpurun_out/r03a_ab_helper_gate.txt
