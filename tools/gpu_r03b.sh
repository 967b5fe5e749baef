#!/bin/bash
# sustained (power-capped) runs with clock-spin and nanosleep staggers; long-context chunking knob at 16384
mkdir -p gpurun_out
{ SUSTAIN_S=3 timeout 600 python tools/sustained.py -350,3,0 -400350,3,0 -400500,3,0 -400250,3,0 -350,3,0 -400350,3,0 -300350,3,0; } 2>&1 | tee gpurun_out/r03b_sustained_stagger_modes.txt
timeout 600 python - <<'PY' 2>&1 | tee gpurun_out/r03b_long_chunk_16384.txt
import torch, json, time, fft_b200
from fft_b200 import _lib
lib = _lib.load()
B, N, C = 16, 16384, 768
V = [torch.randn(B, N, C, device='cuda') for _ in range(2)]
g = [torch.randn(B, C // 16, N // 2 + 1, dtype=torch.cfloat, device='cuda') for _ in range(2)]
alg = fft_b200.plan_info(B, N, N, C, 16)['algorithmic_bytes']
for mb in (0, 24, 48, 96, 0, 48):
    lib.spectre_mix_set_long_chunk_mb(mb)
    for i in range(3): fft_b200.spectral_mix(V[i % 2], g[i % 2], n_fft=N, group_width=16)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(6): fft_b200.spectral_mix(V[i % 2], g[i % 2], n_fft=N, group_width=16)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 6
    print(json.dumps({'long_chunk_mb': mb, 'us': round(ms * 1e3, 1), 'GBps': round(alg / ms / 1e6)}), flush=True)
    time.sleep(0.5)
lib.spectre_mix_set_long_chunk_mb(0)
PY
