#!/usr/bin/env python
"""Interleaved A/B of the host entry's chunk schedules at the metric shape (GPU box): uniform chunks against the ramped schedule."""
import json
import os
import statistics
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fft_b200  # noqa: E402

B, N, C, dg = 148, 4096, 768, 16
hV = torch.randn(B, N, C).pin_memory()
hg = torch.randn(B, C // dg, N // 2 + 1, dtype=torch.cfloat).pin_memory()
ho = torch.empty(B, N, C).pin_memory()
cfgs = {"uniform_32MB": {"SPECTRE_MIX_HOST_CHUNK_MB": "32"}, "uniform_64MB": {"SPECTRE_MIX_HOST_CHUNK_MB": "64"},
        "ramp_12_to_26MB": {"SPECTRE_MIX_HOST_CHUNK_MAX_MB": "26"}, "ramp_12_to_50MB": {"SPECTRE_MIX_HOST_CHUNK_MAX_MB": "50"},
        "ramp_12_to_100MB": {"SPECTRE_MIX_HOST_CHUNK_MAX_MB": "100"}}
res = {k: [] for k in cfgs}
for rnd in range(6):
    for k in (list(cfgs) if rnd % 2 == 0 else list(cfgs)[::-1]):
        for e in ("SPECTRE_MIX_HOST_CHUNK_MB", "SPECTRE_MIX_HOST_CHUNK_MAX_MB"):
            os.environ.pop(e, None)
        os.environ.update(cfgs[k])
        fft_b200.spectral_mix_host(hV, hg, n_fft=N, group_width=dg, out=ho)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            fft_b200.spectral_mix_host(hV, hg, n_fft=N, group_width=dg, out=ho)
        torch.cuda.synchronize()
        res[k].append((time.perf_counter() - t0) / 3 * 1e3)
for k, v in res.items():
    print(json.dumps({"schedule": k, "ms_median": round(statistics.median(v), 2), "ms_min": round(min(v), 2), "ms_all": [round(x, 2) for x in v],
                      "tokens_per_s_median": round(B * N / statistics.median(v) * 1e3)}))
