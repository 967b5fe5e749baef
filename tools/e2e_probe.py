#!/usr/bin/env python
"""Where the host-buffer leg's time goes (GPU box): spectre_mix_fwd_host at several chunk sizes against the SAME chunked copy
pattern without the kernel (4 streams round robin: H2D V, H2D gate, D2H out per chunk) and against two monolithic copies."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fft_b200  # noqa: E402

B, N, C, dg = 148, 4096, 768, 16
NG, FH = C // dg, N // 2 + 1
hV = torch.randn(B, N, C).pin_memory()
hg = torch.randn(B, NG, FH, dtype=torch.cfloat).pin_memory()
ho = torch.empty(B, N, C).pin_memory()
tok = B * N


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


out = {}
for mb in (32, 64):
    os.environ["SPECTRE_MIX_HOST_CHUNK_MB"] = str(mb)
    t = timed(lambda: fft_b200.spectral_mix_host(hV, hg, n_fft=N, group_width=dg, out=ho))
    out["host_entry_chunk_%dMB" % mb] = {"ms": round(t * 1e3, 2), "tokens_per_s": round(tok / t)}
os.environ.pop("SPECTRE_MIX_HOST_CHUNK_MB")
for mx in (0, 50, 76, 100, 152, 202):   # ramped schedule (the default), top of the ramp varied
    if mx:
        os.environ["SPECTRE_MIX_HOST_CHUNK_MAX_MB"] = str(mx)
    t = timed(lambda: fft_b200.spectral_mix_host(hV, hg, n_fft=N, group_width=dg, out=ho))
    out["host_entry_ramped_max_%s" % ("default" if not mx else "%dMB" % mx)] = {"ms": round(t * 1e3, 2), "tokens_per_s": round(tok / t)}
os.environ.pop("SPECTRE_MIX_HOST_CHUNK_MAX_MB", None)

streams = [torch.cuda.Stream() for _ in range(4)]
for rows in (1, 2, 4):
    dv = [torch.empty(rows, N, C, device="cuda") for _ in range(4)]
    dgt = [torch.empty(rows, NG, FH, dtype=torch.cfloat, device="cuda") for _ in range(4)]
    do = [torch.empty(rows, N, C, device="cuda") for _ in range(4)]

    def chunked():
        for k, b0 in enumerate(range(0, B, rows)):
            i = k % 4
            nb = min(rows, B - b0)
            with torch.cuda.stream(streams[i]):
                dv[i][:nb].copy_(hV[b0:b0 + nb], non_blocking=True)
                dgt[i][:nb].copy_(hg[b0:b0 + nb], non_blocking=True)
                ho[b0:b0 + nb].copy_(do[i][:nb], non_blocking=True)
    t = timed(chunked)
    out["copies_only_chunk_%drows" % rows] = {"ms": round(t * 1e3, 2), "tokens_per_s": round(tok / t)}

dV, dO, dG = torch.empty(B, N, C, device="cuda"), torch.empty(B, N, C, device="cuda"), torch.empty(B, NG, FH, dtype=torch.cfloat, device="cuda")
s_up, s_dn = streams[0], streams[1]


def mono(with_gate):
    def f():
        with torch.cuda.stream(s_up):
            dV.copy_(hV, non_blocking=True)
            if with_gate:
                dG.copy_(hg, non_blocking=True)
        with torch.cuda.stream(s_dn):
            ho.copy_(dO, non_blocking=True)
    return f


for wg in (False, True):
    t = timed(mono(wg))
    out["monolithic_copies" + ("_with_gate" if wg else "")] = {"ms": round(t * 1e3, 2), "tokens_per_s": round(tok / t)}
t = timed(lambda: dV.copy_(hV, non_blocking=True))
out["h2d_alone_GBps"] = round(hV.numel() * 4 / t / 1e9, 1)
t = timed(lambda: ho.copy_(dO, non_blocking=True))
out["d2h_alone_GBps"] = round(hV.numel() * 4 / t / 1e9, 1)
print(json.dumps(out, indent=1))
