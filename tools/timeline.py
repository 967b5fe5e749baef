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
if os.environ.get('SPX_ALT'):
    _lib.LIB_PATH = _lib.LIB_PATH.replace('libspectre_mix.so', 'libspectre_mix_alt.so')

ap = argparse.ArgumentParser()
ap.add_argument("--n-fft", type=int, default=4096)
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--prefetch", type=int, default=0)
ap.add_argument("--tma", type=int, default=1)
ap.add_argument("--tmem", type=int, default=1)
ap.add_argument("--skew", type=int, default=0)
ap.add_argument("--sched", type=int, default=0)
a = ap.parse_args()
lib = _lib.load()
lib.spectre_mix_set_prefetch(a.prefetch)
lib.spectre_mix_set_tma(a.tma)
lib.spectre_mix_set_tmem(a.tmem)
lib.spectre_mix_set_skew_ns(a.skew)
lib.spectre_mix_set_sched(a.sched)
dev = torch.device("cuda")
C, dg = 768, 16
V = torch.randn(a.batch, a.n_fft, C, device=dev)
g = torch.randn(a.batch, C // dg, a.n_fft // 2 + 1, dtype=torch.cfloat, device=dev)
for _ in range(2):
    fft_b200.spectral_mix(V, g, n_fft=a.n_fft, group_width=dg)
info = fft_b200.plan_info(a.batch, a.n_fft, a.n_fft, C, dg)
grid = info["grid"]
tl = torch.zeros(grid * 5 * 8 * 8, dtype=torch.int64, device=dev)
lib.spectre_mix_set_timeline(tl.data_ptr())
fft_b200.spectral_mix(V, g, n_fft=a.n_fft, group_width=dg)
torch.cuda.synchronize()
lib.spectre_mix_set_timeline(None)
lib.spectre_mix_set_skew_ns(0)
lib.spectre_mix_set_sched(0)
tall = tl.view(grid, 5, 8, 8).cpu().double()
names = ["wait landing", "F0", "inner fwd", "MID", "inner inv", "I0 math", "output"]
print(f"n_fft={a.n_fft} B={a.batch} grid={grid} skew={a.skew} sched={a.sched} prefetch={a.prefetch} plan={info}")
for grp in range(4):
    t = tall[:, grp]
    if (t[:, 1, 7] > 0).sum() == 0:
        continue
    print(f"-- thread group {grp} (threads {128 * grp}..{128 * grp + 127}); stamps relative to group 0's tile start")
    for tile in range(1, 5):
        ok = (t[:, tile, 7] > 0) & (tall[:, 0, tile, 0] > 0)
        if ok.sum() == 0:
            break
        d = (t[:, tile, 1:] - t[:, tile, :-1])[ok]
        tot = (t[ok, tile, 7] - t[ok, tile, 0])
        gap = (t[ok, tile, 0] - t[ok, tile - 1, 7])
        rel = (t[ok, tile, :] - tall[ok, 0, tile, 0:1]).mean(0)
        print(f"tile#{tile}: total {tot.mean():7.0f} ns  gap-from-prev {gap.mean():5.0f} | " +
              "  ".join(f"{n} {d[:, i].mean():6.0f}" for i, n in enumerate(names)) + " | at " + " ".join(f"{x:6.0f}" for x in rel))

h = tall[:, 4]
if (h[:, 3, 2] > 0).sum() > 0:
    print("-- helper warpgroup (elected thread; diagnostic build -DSPX_HELPER_TL=1): per phase, ns / cycles summed over 16 steps")
    for P in range(2, 7):
        ok = h[:, P, 2] > 0
        if ok.sum() == 0:
            break
        x = h[ok, P]
        rel0 = (x[:, 0] - tall[ok, 0, max(P - 1, 0), 0]).mean()
        print(f"phase {P}: entry at {rel0:6.0f} ns after compute tile#{P-1} start | wait for compute {(x[:,1]-x[:,0]).mean():6.0f} ns  steps {(x[:,2]-x[:,1]).mean():6.0f} ns | "
              f"cycles: wait landed {x[:,3].mean():6.0f}  move {x[:,4].mean():6.0f}  barrier {x[:,5].mean():6.0f}  tma duties {x[:,6].mean():6.0f}")
