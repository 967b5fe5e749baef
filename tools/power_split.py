#!/usr/bin/env python
"""Board power and SM clock of the n_fft = 4096 kernel's halves under continuous load (GPU box): the full kernel, the FFT passes
without tile I/O (sched bit 3), tile I/O without FFT passes (sched bit 2) and a plain device copy, a few seconds each, power and
clock sampled by NVML every 50 ms during the run.  Diagnostic modes give invalid results; this is an energy inventory."""
import ctypes
import json
import os
import statistics
import sys
import threading
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fft_b200 import _lib  # noqa: E402
import pynvml  # noqa: E402

if os.environ.get("SPX_ALT"):
    _lib.LIB_PATH = _lib.LIB_PATH.replace("libspectre_mix.so", "libspectre_mix_%s.so" % os.environ["SPX_ALT"])
lib = _lib.load()
lib.spectre_mix_set_prefetch(int(os.environ.get("PS_PREFETCH", "0")))
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
n_fft, C, dg, B = 4096, 768, 16, 148
dev = torch.device("cuda")
gen = torch.Generator(device=dev).manual_seed(0)
V = [torch.randn(B, n_fft, C, device=dev, generator=gen) for _ in range(2)]
g = [torch.randn(B, C // dg, n_fft // 2 + 1, dtype=torch.cfloat, device=dev, generator=gen) for _ in range(2)]
o = [torch.empty(B, n_fft, C, device=dev) for _ in range(2)]
st = torch.cuda.current_stream().cuda_stream
alg = B * n_fft * C * 8 + B * (C // dg) * (n_fft // 2 + 1) * 8
tiles_per_sm = B * (C // 8) / 148.0


def mix(i):
    i %= 2
    lib.spectre_mix_fwd(V[i].data_ptr(), 0, V[i].stride(0), V[i].stride(1), g[i].data_ptr(), None, C, o[i].data_ptr(), 0,
                        o[i].stride(0), o[i].stride(1), B, n_fft, n_fft, C, dg, ctypes.c_void_p(st))


def copy(i):
    o[i % 2].copy_(V[i % 2])


def sustained(fn, seconds, per_call_s):
    samples = []
    stop = threading.Event()

    def sampler():
        while not stop.is_set():
            samples.append((time.perf_counter(), pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0, pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
            time.sleep(0.05)
    th = threading.Thread(target=sampler, daemon=True)
    n = max(30, int(seconds / per_call_s))
    third = n // 3
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    th.start()
    t0 = time.perf_counter()
    evs[0].record()
    for k in range(3):
        for i in range(third):
            fn(i)
            if i % 64 == 63:
                evs[k].query()   # keep the launch queue bounded without a sync
        evs[k + 1].record()
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    stop.set()
    th.join()
    late = [s for s in samples if t0 + 0.6 * (t1 - t0) <= s[0] <= t1 - 0.05]
    ms = [evs[k].elapsed_time(evs[k + 1]) / third for k in range(3)]
    return ms, (round(statistics.median(s[1] for s in late)) if late else None), (round(statistics.median(s[2] for s in late)) if late else None)


secs = float(os.environ.get("SUSTAIN_S", "3"))
modes = (("full kernel", 3, mix, 1e-3), ("tile I/O only (no FFT passes)", 3 | 4, mix, 0.8e-3), ("full kernel again", 3, mix, 1e-3)) if os.environ.get("PS_SHORT") else None
for name, sched, fn, per in modes or (("full kernel", 3, mix, 1e-3), ("FFT passes only (no tile I/O)", 3 | 8, mix, 0.8e-3), ("tile I/O only (no FFT passes)", 3 | 4, mix, 0.8e-3),
                             ("torch copy of the same tensor", 3, copy, 0.6e-3), ("full kernel again", 3, mix, 1e-3)):
    lib.spectre_mix_set_sched(sched)
    time.sleep(3.0)
    ms, pw, clk = sustained(fn, secs, per)
    rec = {"what": name, "ms_per_launch_by_third": [round(x, 4) for x in ms], "us_per_tile_last_third": round(ms[2] * 1e3 / tiles_per_sm, 2),
           "power_w_median_late": pw, "sm_mhz_median_late": clk}
    if fn is mix:
        rec["GBps_last_third"] = round(alg / ms[2] / 1e6)
    else:
        rec["GBps_last_third"] = round(2 * V[0].numel() * 4 / ms[2] / 1e6)
    rec["joules_per_launch"] = round(pw * ms[2] * 1e-3, 3) if pw else None
    print(json.dumps(rec), flush=True)
lib.spectre_mix_set_sched(3)
lib.spectre_mix_set_prefetch(0)
