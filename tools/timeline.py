#!/usr/bin/env python
"""Per-phase timeline of the mix kernel from in-kernel %globaltimer stamps (GPU box).

slots: 0 tile start, 1 tile landed (TMA), 2 F0 done, 3 inner fwd done, 4 MID done, 5 inner inv done (+gate put),
       6 I0 butterflies done, 7 outputs handed to TMA / stored
"""
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
ap.add_argument("--prefetch", type=int, default=1)
ap.add_argument("--tma", type=int, default=1)
ap.add_argument("--tmem", type=int, default=1)
a = ap.parse_args()
lib = _lib.load()
lib.spectre_mix_set_prefetch(a.prefetch)
lib.spectre_mix_set_tma(a.tma)
lib.spectre_mix_set_tmem(a.tmem)
dev = torch.device("cuda")
C, dg = 768, 16
V = torch.randn(a.batch, a.n_fft, C, device=dev)
g = torch.randn(a.batch, C // dg, a.n_fft // 2 + 1, dtype=torch.cfloat, device=dev)
for _ in range(2):
    fft_b200.spectral_mix(V, g, n_fft=a.n_fft, group_width=dg)
info = fft_b200.plan_info(a.batch, a.n_fft, a.n_fft, C, dg)
grid = info["grid"]
tl = torch.zeros(grid * 8 * 8, dtype=torch.int64, device=dev)
lib.spectre_mix_set_timeline(tl.data_ptr())
fft_b200.spectral_mix(V, g, n_fft=a.n_fft, group_width=dg)
torch.cuda.synchronize()
lib.spectre_mix_set_timeline(None)
t = tl.view(grid, 8, 8).cpu().double()
names = ["wait landing", "F0", "inner fwd", "MID", "inner inv", "I0 math", "output"]
print(f"n_fft={a.n_fft} B={a.batch} grid={grid} plan={info}")
for tile in range(1, 6):
    d = (t[:, tile, 1:] - t[:, tile, :-1])
    ok = t[:, tile, 7] > 0
    if ok.sum() == 0:
        break
    d = d[ok]
    tot = (t[ok, tile, 7] - t[ok, tile, 0])
    gap = (t[ok, tile, 0] - t[ok, tile - 1, 7])
    print(f"tile#{tile}: total {tot.mean():7.0f} ns (min {tot.min():.0f} max {tot.max():.0f})  gap-from-prev {gap.mean():5.0f} | " +
          "  ".join(f"{n} {d[:, i].mean():6.0f}" for i, n in enumerate(names)))
