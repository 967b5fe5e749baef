#!/usr/bin/env python
"""Run a few decode steps (update + pruned read-out) on one configuration (for ncu captures)."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fft_b200  # noqa: E402
from fft_b200 import _lib  # noqa: E402

lib = _lib.load()
dev = torch.device("cuda")
n, d, G = 4096, 768, 48
cache = fft_b200.PrefixFFTCache(n, d, device=dev)
cache.prefill(torch.randn(4000, d, device=dev), torch.randn(4000, d, device=dev))
gate = torch.randn(G, n // 2 + 1, dtype=torch.cfloat, device=dev)
v, vb, out = torch.randn(d, device=dev), torch.randn(d, device=dev), torch.empty(d, device=dev)
for _ in range(4):
    rc = lib.spectre_decode_step(cache.prefix_fft.data_ptr(), v.data_ptr(), vb.data_ptr(), gate.data_ptr(), out.data_ptr(), n, d, d // G, 5000,
                                 cache._ws.data_ptr(), cache._ws_bytes, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0
torch.cuda.synchronize()
print("ok", float(out.abs().mean()))
