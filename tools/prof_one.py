#!/usr/bin/env python
"""Launch the mix kernel a few times on one configuration (for ncu captures)."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fft_b200  # noqa: E402
from fft_b200 import _lib  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n-fft", type=int, default=4096)
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--channels", type=int, default=768)
ap.add_argument("--group-width", type=int, default=16)
ap.add_argument("--tile", type=int, default=0)
ap.add_argument("--prefetch", type=int, default=0)
ap.add_argument("--reps", type=int, default=4)
ap.add_argument("--sched", type=int, default=None, help="scheduling / diagnostic flags (4: tile I/O only, 8: FFT passes only)")
ap.add_argument("--bf16", action="store_true")
ap.add_argument("--mem", action="store_true")
a = ap.parse_args()
lib = _lib.load()
lib.spectre_mix_set_tile_channels(a.tile)
lib.spectre_mix_set_prefetch(a.prefetch)
if a.sched is not None:
    lib.spectre_mix_set_sched(a.sched)
dev = torch.device("cuda")
V = torch.randn(a.batch, a.n_fft, a.channels, device=dev)
if a.bf16:
    V = V.bfloat16()
g = torch.randn(a.batch, a.channels // a.group_width, a.n_fft // 2 + 1, dtype=torch.cfloat, device=dev)
m = torch.randn(a.n_fft // 2 + 1, a.channels, dtype=torch.cfloat, device=dev) if a.mem else None
for _ in range(a.reps):
    y = fft_b200.spectral_mix(V, g, m, n_fft=a.n_fft, group_width=a.group_width)
torch.cuda.synchronize()
print("ok", float(y.float().abs().mean()))
