#!/bin/bash
mkdir -p gpurun_out
for alt in ga3 ga3x1 ga3x3; do
SPX_ALT=$alt timeout 300 python - <<'PY'
import os, torch, fft_b200
from fft_b200 import _lib
_lib.LIB_PATH = _lib.LIB_PATH.replace("libspectre_mix.so", "libspectre_mix_%s.so" % os.environ["SPX_ALT"])
torch.manual_seed(1)
for (B, N, C, mem) in [(3, 4096, 768, True), (5, 3000, 72, False), (150, 4096, 64, False)]:
    V = torch.randn(B, N, C, device='cuda'); g = torch.randn(B, C // 8, 2049, dtype=torch.cfloat, device='cuda')
    m = torch.randn(2049, C, dtype=torch.cfloat, device='cuda') / 8 if mem else None
    y = fft_b200.spectral_mix(V, g, m, n_fft=4096, group_width=8)
    Vf = torch.fft.rfft(V.double(), n=4096, dim=1); gb = g.to(torch.complex128).permute(0, 2, 1).repeat_interleave(8, -1)
    want = torch.fft.irfft(gb * Vf + (m.to(torch.complex128) if mem else 0), n=4096, dim=1)[:, :N].float()
    print(os.environ["SPX_ALT"], (B, N, C, mem), 'rel-L2', float((y - want).norm() / want.norm()))
PY
done 2>&1 | tee gpurun_out/r03i_parity_alt.txt
{ for alt in "" ga3 ga3x1 ga3x3 "" ga3 ga3x1; do echo "== alt='$alt'"; SPX_ALT=$alt AB_BATCH=148 timeout 300 python tools/ab.py -350,3,0 -500,3,0; done; } 2>&1 | tee gpurun_out/r03i_ab_raw_async_gate.txt
