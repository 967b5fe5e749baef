#!/bin/bash
# session r04p: host-entry release test; L2 evict-first hints on the TMA loads / stores (SPX_L2HINT=1|2|3): burst A/B and sustained energy
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "host_entry" 2>&1 | tail -2
{ for rep in 1 2; do for alt in "" l2h1 l2h2 l2h3; do echo "== alt='$alt'"; SPX_ALT=$alt AB_BATCH=148 timeout 300 python tools/ab.py -350,3,0; done; done
  for alt in "" l2h1 l2h2 l2h3; do echo "== sustained alt='$alt'"; SPX_ALT=$alt PS_SHORT=1 timeout 300 python tools/power_split.py | grep -v "tile I/O"; done; } 2>&1 | tee gpurun_out/r04p_ab_l2hint.txt
