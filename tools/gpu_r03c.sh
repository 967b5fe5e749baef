#!/bin/bash
mkdir -p gpurun_out
timeout 600 python - <<'PY' 2>&1 | tee gpurun_out/r03c_long_chunk_16384.txt
import os, torch, json, time, fft_b200
B, N = 16, 16384
for C in (768, 256):
    V = [torch.randn(B, N, C, device='cuda') for _ in range(2)]
    g = [torch.randn(B, C // 16, N // 2 + 1, dtype=torch.cfloat, device='cuda') for _ in range(2)]
    alg = fft_b200.plan_info(B, N, N, C, 16)['algorithmic_bytes']
    for mb in (0, 1, 100, 0, 1):
        os.environ['SPECTRE_MIX_LONG_CHUNK_MB'] = str(mb)
        for i in range(3): fft_b200.spectral_mix(V[i % 2], g[i % 2], n_fft=N, group_width=16)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(6): fft_b200.spectral_mix(V[i % 2], g[i % 2], n_fft=N, group_width=16)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 6
        print(json.dumps({'C': C, 'long_chunk_mb': mb, 'rows_per_chunk': 'all' if mb == 0 else max(1, (mb << 20) // (N * C * 4)), 'us': round(ms * 1e3, 1), 'GBps': round(alg / ms / 1e6)}), flush=True)
        time.sleep(0.5)
PY
