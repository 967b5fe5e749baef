#!/bin/bash
# L2 promotion of the input tensor map x n_fft (kernel-only), then the e2e leg of the bench twice
mkdir -p gpurun_out
python - <<'PY' 2>&1 | tee gpurun_out/promo.log
import sys, os, json, torch
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tools"))
from tune import time_case
from fft_b200 import _lib
lib = _lib.load()
for n_fft, B in ((4096, 128), (1024, 512), (2048, 256), (256, 2048)):
    for promo in (0, 1, 2, 3, 0):
        lib.spectre_mix_set_l2_promotion(promo)
        r = time_case(lib, n_fft, 768, 16, B, 0, 0, tma=1, tmem=1, reps=40)
        print(json.dumps(dict(n_fft=n_fft, promo=promo, GBps=round(r["GBps"]), ms=round(r["ms"], 4))), flush=True)
lib.spectre_mix_set_l2_promotion(0)
PY
for i in 1 2; do timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu 2>/dev/null | tail -1 | python -c "import sys, json; d = json.loads(sys.stdin.read()); print('bench value %.4g frac %.3f e2e %.4g' % (d['value'], d['roofline']['frac'], d['e2e']['value']))"; done
