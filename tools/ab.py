#!/usr/bin/env python
"""Round-robin A/B timing of launch-time knobs of the n_fft = 4096 kernel (GPU box).

Sustained back-to-back launches pull the SM clock down (power cap), so configurations timed one after the other are not
comparable; this harness allocates once, cycles through the configurations several times with short bursts, and reports
the median / best burst per configuration together with the SM clock read right after each burst (NVML).

    python tools/ab.py "skew,sched,prefetch" ...      e.g.  python tools/ab.py -300,3,0 0,0,0 -300,11,0
"""
import ctypes
import json
import os
import statistics
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fft_b200 import _lib  # noqa: E402

if os.environ.get("SPX_ALT"):
    _lib.LIB_PATH = _lib.LIB_PATH.replace("libspectre_mix.so", "libspectre_mix_%s.so" % os.environ["SPX_ALT"])
lib = _lib.load()
try:
    import pynvml
    pynvml.nvmlInit()
    _h = pynvml.nvmlDeviceGetHandleByIndex(0)

    def sm_clock():
        return pynvml.nvmlDeviceGetClockInfo(_h, pynvml.NVML_CLOCK_SM)
except Exception:  # noqa: BLE001
    def sm_clock():
        return -1

n_fft, C, dg = int(os.environ.get("AB_NFFT", "4096")), 768, 16
B = int(os.environ.get("AB_BATCH", "128"))
rounds = int(os.environ.get("AB_ROUNDS", "5"))
burst = int(os.environ.get("AB_BURST", "8"))
dev = torch.device("cuda")
dtype = torch.bfloat16 if os.environ.get("AB_DTYPE") == "bf16" else torch.float32
dt = 1 if dtype == torch.bfloat16 else 0
lib.spectre_mix_set_tile_channels(int(os.environ.get("AB_TILE", "0")))
gen = torch.Generator(device=dev).manual_seed(0)
V = [torch.randn(B, n_fft, C, device=dev, generator=gen).to(dtype) for _ in range(2)]
g = [torch.randn(B, C // dg, n_fft // 2 + 1, dtype=torch.cfloat, device=dev, generator=gen) for _ in range(2)]
o = [torch.empty(B, n_fft, C, device=dev, dtype=dtype) for _ in range(2)]
st = torch.cuda.current_stream().cuda_stream
alg = B * n_fft * C * (4 if dt else 8) + B * (C // dg) * (n_fft // 2 + 1) * 8
tiles = B * (C // 8) / 148.0   # 8-channel tiles per SM (meaningful for the 4096 plan)


def run(i):
    i %= 2
    rc = lib.spectre_mix_fwd(V[i].data_ptr(), dt, V[i].stride(0), V[i].stride(1), g[i].data_ptr(), None, C, o[i].data_ptr(), dt,
                             o[i].stride(0), o[i].stride(1), B, n_fft, n_fft, C, dg, ctypes.c_void_p(st))
    if rc:
        raise RuntimeError(lib.spectre_mix_last_error().decode())


cfgs = [tuple(int(x) for x in a.split(",")) for a in sys.argv[1:]] or [(-300, 3, 0)]
res = {c: [] for c in cfgs}
clk = {c: [] for c in cfgs}
for i in range(3):
    run(i)
torch.cuda.synchronize()
for r in range(rounds):
    for c in (cfgs if r % 2 == 0 else cfgs[::-1]):
        lib.spectre_mix_set_skew_ns(c[0]); lib.spectre_mix_set_sched(c[1]); lib.spectre_mix_set_prefetch(c[2])
        run(0)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(burst):
            run(i)
        e1.record()
        torch.cuda.synchronize()
        clk[c].append(sm_clock())
        res[c].append(e0.elapsed_time(e1) / burst)
        time.sleep(0.15)
lib.spectre_mix_set_skew_ns(-350); lib.spectre_mix_set_sched(3); lib.spectre_mix_set_prefetch(0)
for c in cfgs:
    ms = statistics.median(res[c])
    print(json.dumps(dict(skew=c[0], sched=c[1], prefetch=c[2], GBps_median=round(alg / ms / 1e6), GBps_best=round(alg / min(res[c]) / 1e6),
                          us_per_tile=round(ms * 1e3 / tiles, 2), sm_mhz=clk[c])), flush=True)
