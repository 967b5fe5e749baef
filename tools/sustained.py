#!/usr/bin/env python
"""Sustained (power-capped) throughput of the n_fft = 4096 kernel: each configuration runs back to back for a few seconds and
the rate of the last third is reported with the SM clock and power read at the end (GPU box).

    python tools/sustained.py "skew,sched,prefetch" ...
"""
import ctypes
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fft_b200 import _lib  # noqa: E402

if os.environ.get("SPX_ALT"):
    _lib.LIB_PATH = _lib.LIB_PATH.replace("libspectre_mix.so", "libspectre_mix_%s.so" % os.environ["SPX_ALT"])
lib = _lib.load()
import pynvml  # noqa: E402

pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
n_fft, C, dg, B = 4096, 768, 16, 128
dev = torch.device("cuda")
gen = torch.Generator(device=dev).manual_seed(0)
V = [torch.randn(B, n_fft, C, device=dev, generator=gen) for _ in range(2)]
g = [torch.randn(B, C // dg, n_fft // 2 + 1, dtype=torch.cfloat, device=dev, generator=gen) for _ in range(2)]
o = [torch.empty(B, n_fft, C, device=dev) for _ in range(2)]
st = torch.cuda.current_stream().cuda_stream
alg = B * n_fft * C * 8 + B * (C // dg) * (n_fft // 2 + 1) * 8


def run(i):
    i %= 2
    lib.spectre_mix_fwd(V[i].data_ptr(), 0, V[i].stride(0), V[i].stride(1), g[i].data_ptr(), None, C, o[i].data_ptr(), 0,
                        o[i].stride(0), o[i].stride(1), B, n_fft, n_fft, C, dg, ctypes.c_void_p(st))


seconds = float(os.environ.get("SUSTAIN_S", "3"))
for a in sys.argv[1:]:
    c = tuple(int(x) for x in a.split(","))
    lib.spectre_mix_set_skew_ns(c[0]); lib.spectre_mix_set_sched(c[1]); lib.spectre_mix_set_prefetch(c[2])
    time.sleep(2.0)                                   # cool down between configurations
    n = int(seconds / 0.8e-3)
    third = n // 3
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    evs[0].record()
    for k in range(3):
        for i in range(third):
            run(i)
        evs[k + 1].record()
    torch.cuda.synchronize()
    clk = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
    pw = pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0
    rates = [round(alg * third / (evs[k].elapsed_time(evs[k + 1]) * 1e6)) for k in range(3)]
    print(json.dumps(dict(skew=c[0], sched=c[1], prefetch=c[2], GBps_by_third=rates, sm_mhz_end=clk, power_w_end=round(pw))), flush=True)
lib.spectre_mix_set_skew_ns(-350); lib.spectre_mix_set_sched(3); lib.spectre_mix_set_prefetch(0)
