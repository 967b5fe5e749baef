#!/bin/bash
python - <<'PY'
import sys, os, json, torch
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tools"))
from tune import time_case
from fft_b200 import _lib
lib = _lib.load()
for skew in (0, 200, 400, 600, 800, 1000, 1400):
    lib.spectre_mix_set_skew_ns(skew)
    for tmem in (1, 0):
        r = time_case(lib, 4096, 768, 16, 128, 0, 1, tma=1, tmem=tmem, reps=20)
        print(json.dumps(dict(skew=skew, tmem=tmem, GBps=round(r["GBps"]), ms=round(r["ms"], 4))), flush=True)
lib.spectre_mix_set_skew_ns(0)
PY
