#!/usr/bin/env python
"""Race hunt for the TMEM-staged n_fft = 4096 variants (GPU box): random batch sizes, channel counts, row counts and dtypes,
default schedule vs the plain TMA variant (tmem off, no stagger) -- the two must agree bit for bit on every trial."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fft_b200  # noqa: E402
from fft_b200 import _lib  # noqa: E402

lib = _lib.load()
dev = torch.device("cuda")
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
trials = int(sys.argv[1]) if len(sys.argv) > 1 else 150
bad = 0
for t in range(trials):
    B = int(rng.integers(1, 70))
    dg = int(rng.choice([4, 8, 12, 16]))
    C = dg * int(rng.integers(1, 768 // dg + 1))
    N = 4096 if rng.random() < 0.6 else int(rng.integers(1, 5000))
    bf16 = rng.random() < 0.3
    mem = rng.random() < 0.25
    g = torch.Generator(device=dev).manual_seed(t)
    V = torch.randn(B, N, C, device=dev, generator=g)
    if bf16:
        V = V.bfloat16()
    gate = torch.randn(B, C // dg, 2049, dtype=torch.cfloat, device=dev, generator=g)
    m = torch.randn(2049, C, dtype=torch.cfloat, device=dev, generator=g) if mem else None
    ys = [fft_b200.spectral_mix(V, gate, m, n_fft=4096, group_width=dg) for _ in range(2)]
    lib.spectre_mix_set_tmem(0); lib.spectre_mix_set_skew_ns(0); lib.spectre_mix_set_sched(0)
    ref = fft_b200.spectral_mix(V, gate, m, n_fft=4096, group_width=dg)
    lib.spectre_mix_set_tmem(1); lib.spectre_mix_set_skew_ns(-350); lib.spectre_mix_set_sched(3)
    ok = all(torch.equal(y, ref) for y in ys) and bool(torch.isfinite(ref.float()).all())
    if not ok:
        bad += 1
        print("MISMATCH", dict(trial=t, B=B, C=C, dg=dg, N=N, bf16=bf16, mem=mem), flush=True)
print(f"{trials} trials, {bad} mismatches")
sys.exit(1 if bad else 0)
