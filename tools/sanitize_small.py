#!/usr/bin/env python
"""Small run of every kernel family for compute-sanitizer (GPU box):
    compute-sanitizer --tool memcheck python tools/sanitize_small.py
Checks results against the unfused / oracle-free compositions only loosely (finite, right shape); parity is the tests' job."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fft_b200  # noqa: E402
from fft_b200 import ops  # noqa: E402

dev = torch.device("cuda")
torch.manual_seed(0)


def case(B, N, n_fft, C, dg, dt=torch.float32, mem=False):
    V = torch.randn(B, N, C, device=dev).to(dt)
    g = torch.randn(B, C // dg, n_fft // 2 + 1, dtype=torch.cfloat, device=dev)
    m = torch.randn(n_fft // 2 + 1, C, dtype=torch.cfloat, device=dev) if mem else None
    y = fft_b200.spectral_mix(V, g, m, n_fft=n_fft, group_width=dg)
    assert torch.isfinite(y.float()).all()
    return V, g


for args in [(2, 64, 64, 16, 4), (2, 1000, 1024, 32, 16), (1, 2048, 2048, 32, 8), (3, 4096, 4096, 32, 16), (2, 4000, 4096, 24, 8),
             (2, 4096, 4096, 32, 16, torch.bfloat16), (1, 8192, 8192, 32, 16), (1, 16384, 16384, 16, 16), (2, 4096, 4096, 12, 6),
             (2, 4096, 4096, 9, 3), (2, 4096, 4096, 32, 16, torch.float32, True),
             # wide-row TMEM variants (batch large enough to select them), DIT2 at 8192 (even rows) and its odd-row fallback
             (40, 1024, 1024, 800, 16), (24, 2000, 2048, 784, 16, torch.float32, True), (2, 8190, 8192, 32, 16, torch.float32, True),
             (1, 8191, 8192, 32, 16)]:
    case(*args)
# fused gate generator
B, n_fft, G, dg = 2, 4096, 4, 16
a = torch.randn(B, 2 * G, 45, dtype=torch.cfloat, device=dev)
bias = torch.randn(2 * G, n_fft // 2 + 1, device=dev) * 0.3
eps = torch.full((2 * G,), 1e-4, device=dev)
V = torch.randn(B, n_fft, 2 * G * dg, device=dev)
y = fft_b200.spectral_mix_anchors(V, a, bias, eps, n_fft=n_fft, group_width=dg, G=G)
assert torch.isfinite(y).all()
# gate gradient
dgate = ops._dgate_fused(V, torch.randn_like(V), n_fft, dg)
assert dgate is not None and torch.isfinite(torch.view_as_real(dgate)).all()
# decode
cache = fft_b200.PrefixFFTCache(256, 32, device=dev)
cache.prefill(torch.randn(200, 32, device=dev), torch.randn(200, 32, device=dev))
gate = torch.randn(4, 129, dtype=torch.cfloat, device=dev)
for _ in range(70):
    q, v = torch.randn(32, device=dev), torch.randn(32, device=dev)
    j = cache._advance(q)
    o = cache.fused_step(v, cache.V_buf[j], gate)
    cache._store_v(j, v)
assert torch.isfinite(o).all()
torch.cuda.synchronize()
print("sanitize_small ok")
