#!/bin/bash
# quick perf check of the flagship shapes (+ optional ncu capture with tag $1)
mkdir -p gpurun_out
python - <<'PY'
import sys, os, json, torch
sys.path.insert(0, os.getcwd())
sys.path.insert(0, os.path.join(os.getcwd(), "tools"))
from tune import time_case
from fft_b200 import _lib
lib = _lib.load()
for n_fft, B in [(4096, 64), (4096, 256), (1024, 256), (2048, 128), (256, 1024)]:
    for tmem, pf in (((1, 1), (1, 0), (0, 0)) if n_fft == 4096 else ((0, 0),)):
        r = time_case(lib, n_fft, 768, 16, B, 0, pf, tma=1, tmem=tmem, reps=20)
        r.update(n_fft=n_fft, B=B, tmem=tmem, prefetch=pf)
        print(json.dumps(r), flush=True)
lib.spectre_mix_set_tmem(1)
PY
if [ -n "$1" ]; then
  TAG=$1; shift
  bash tools/gpu_prof.sh $TAG "$@"
fi
