#!/bin/bash
# Decomposition of the n_fft = 4096 kernel: full / tile I/O only (sched 4) / FFT passes only (sched 8); results of 4 and 8 are invalid by design
mkdir -p gpurun_out
rm -f gpurun_out/diag.log
for ALT in ""; do SPX_ALT=$ALT python - <<'PY' 2>&1 | tee -a gpurun_out/diag.log
import sys, os, json, torch
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tools"))
from tune import time_case
from fft_b200 import _lib
if os.environ.get('SPX_ALT'):
    _lib.LIB_PATH = _lib.LIB_PATH.replace('libspectre_mix.so', 'libspectre_mix_alt.so')
    print('ALT library', _lib.LIB_PATH)
import fft_b200
lib = _lib.load()
dev = torch.device("cuda")
torch.manual_seed(0)
V = torch.randn(40, 4096, 768, device=dev)
g = torch.randn(40, 48, 2049, dtype=torch.cfloat, device=dev)
lib.spectre_mix_set_skew_ns(0); lib.spectre_mix_set_sched(0)
ref = fft_b200.spectral_mix(V, g, None, n_fft=4096, group_width=16).clone()
def check(skew, sched, pf):
    lib.spectre_mix_set_skew_ns(skew); lib.spectre_mix_set_sched(sched); lib.spectre_mix_set_prefetch(pf)
    ok = True
    for _ in range(2):
        ok &= bool(torch.equal(fft_b200.spectral_mix(V, g, None, n_fft=4096, group_width=16), ref))
    return ok
for skew, sched in [(-300, 3), (-200, 3), (-250, 3), (-300, 3), (-350, 3), (-400, 3), (-100300, 3), (-100500, 3), (-200300, 3), (-200500, 3), (-300, 3), (-300, 2), (-250, 2)]:
    for pf in (0,):
        ok = check(skew, sched, pf) if not (sched & 12) else None
        lib.spectre_mix_set_skew_ns(skew); lib.spectre_mix_set_sched(sched)
        r = time_case(lib, 4096, 768, 16, 128, 0, pf, tma=1, tmem=1, reps=40)
        tiles = 128 * 96 / 148.0
        print(json.dumps(dict(skew=skew, sched=sched, prefetch=pf, exact=ok, GBps=round(r["GBps"]), ms=round(r["ms"], 4), us_per_tile=round(r["ms"] * 1e3 / tiles, 2))), flush=True)
lib.spectre_mix_set_skew_ns(-350); lib.spectre_mix_set_sched(3); lib.spectre_mix_set_prefetch(0)
PY
done
