#!/bin/bash
# session r04a: SPX_TW4 (4 stored stage-0 twiddle rows, 8 ring slots) -- parity, then A/B against the default build
mkdir -p gpurun_out
SPX_ALT=tw4 timeout 300 python - <<'PY' 2>&1 | tee gpurun_out/r04a_parity_tw4.txt
import os, torch, fft_b200
from fft_b200 import _lib
_lib.LIB_PATH = _lib.LIB_PATH.replace("libspectre_mix.so", "libspectre_mix_%s.so" % os.environ["SPX_ALT"])
torch.manual_seed(1)
for (B, N, nf, C, mem) in [(3, 4096, 4096, 768, True), (5, 3000, 4096, 72, False), (150, 4096, 4096, 64, False), (3, 8192, 8192, 768, False),
                           (2, 16384, 16384, 64, False), (4, 1024, 1024, 768, False), (4, 2048, 2048, 768, True)]:
    V = torch.randn(B, N, C, device='cuda'); g = torch.randn(B, C // 8, nf // 2 + 1, dtype=torch.cfloat, device='cuda')
    m = torch.randn(nf // 2 + 1, C, dtype=torch.cfloat, device='cuda') / 8 if mem else None
    y = fft_b200.spectral_mix(V, g, m, n_fft=nf, group_width=8)
    Vf = torch.fft.rfft(V.double(), n=nf, dim=1); gb = g.to(torch.complex128).permute(0, 2, 1).repeat_interleave(8, -1)
    want = torch.fft.irfft(gb * Vf + (m.to(torch.complex128) if mem else 0), n=nf, dim=1)[:, :N].float()
    print(os.environ["SPX_ALT"], (B, N, nf, C, mem), 'rel-L2', float((y - want).norm() / want.norm()))
PY
{ for alt in "" tw4 "" tw4 "" tw4; do echo "== alt='$alt'"; SPX_ALT=$alt AB_BATCH=148 timeout 300 python tools/ab.py -350,3,0 -500,3,0 -200,3,0; done; } 2>&1 | tee gpurun_out/r04a_ab_tw4.txt
