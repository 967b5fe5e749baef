#!/usr/bin/env python
"""Fused gate generator (spectre_mix_fwd_anchors) vs gate_expand + spectre_mix_fwd at the metric shape (GPU box).
Round-robin bursts; prints microseconds per call for: mix alone (materialised gate given), gate_expand + mix, fused."""
import json
import os
import statistics
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fft_b200  # noqa: E402

from fft_b200 import _lib  # noqa: E402
_lib.load().spectre_mix_set_sched(int(os.environ.get("AB_SCHED", "3")))
dev = torch.device("cuda")
B = int(os.environ.get("AB_BATCH", "148"))
n_fft, C, dg, G, H = int(os.environ.get("AB_NFFT", "4096")), 768, 16, 4, 12
F_half, NG = n_fft // 2 + 1, C // dg
Bk = max(4, int(F_half ** 0.5))
gen = torch.Generator(device=dev).manual_seed(0)
V = [torch.randn(B, n_fft, C, device=dev, generator=gen) for _ in range(2)]
a = [torch.randn(B, NG, Bk, dtype=torch.cfloat, device=dev, generator=gen) for _ in range(2)]
bias = 0.3 * torch.randn(NG, F_half, device=dev, generator=gen) - 0.1
eps = torch.full((NG,), 1e-4, device=dev)
gate = [fft_b200.gate_expand(x, bias, eps, None, F_half=F_half, G=G) for x in a]


def mix(i):
    return fft_b200.spectral_mix(V[i], gate[i], n_fft=n_fft, group_width=dg)


def unfused(i):
    return fft_b200.spectral_mix(V[i], fft_b200.gate_expand(a[i], bias, eps, None, F_half=F_half, G=G), n_fft=n_fft, group_width=dg)


def fused(i):
    return fft_b200.spectral_mix_anchors(V[i], a[i], bias, eps, n_fft=n_fft, group_width=dg, G=G)


fns = {"mix_only": mix, "gate_expand+mix": unfused, "fused_anchors": fused}
res = {k: [] for k in fns}
for f in fns.values():
    f(0)
torch.cuda.synchronize()
for r in range(5):
    for k, f in fns.items():
        f(1)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(8):
            f(i % 2)
        e1.record()
        torch.cuda.synchronize()
        res[k].append(e0.elapsed_time(e1) / 8 * 1e3)
        time.sleep(0.2)
err = float((fused(0) - unfused(0)).norm() / unfused(0).norm())
print(json.dumps({"sched": int(os.environ.get("AB_SCHED", "3")), "B": B, "n_fft": n_fft, "us_per_call_median": {k: round(statistics.median(v), 1) for k, v in res.items()},
                  "us_per_call_best": {k: round(min(v), 1) for k, v in res.items()}, "fused_vs_unfused_rel_l2": err}))
