#!/usr/bin/env python
"""Kernel-only timing sweep (GPU box): tile width x prefetch x n_fft, CUDA events, inputs larger than L2.

    python tools/tune.py [--out gpurun_out/tune.json]
"""
import argparse
import ctypes
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fft_b200 import _lib  # noqa: E402


def time_case(lib, n_fft, C, dg, B, tile, prefetch, dtype=torch.float32, mem=False, reps=10, N=None, tma=1, tmem=1):
    dev = torch.device("cuda")
    N = N or n_fft
    lib.spectre_mix_set_tile_channels(tile)
    lib.spectre_mix_set_prefetch(prefetch)
    lib.spectre_mix_set_tma(tma)
    lib.spectre_mix_set_tmem(tmem)
    gen = torch.Generator(device=dev).manual_seed(0)
    sets = 2
    V = [torch.randn(B, N, C, device=dev, generator=gen).to(dtype) for _ in range(sets)]
    g = [torch.randn(B, C // dg, n_fft // 2 + 1, dtype=torch.cfloat, device=dev, generator=gen) for _ in range(sets)]
    o = [torch.empty(B, min(N, n_fft), C, device=dev, dtype=dtype) for _ in range(sets)]
    m = torch.randn(n_fft // 2 + 1, C, dtype=torch.cfloat, device=dev, generator=gen) if mem else None
    st = torch.cuda.current_stream().cuda_stream
    dt = 0 if dtype == torch.float32 else 1

    def run(i):
        i %= sets
        rc = lib.spectre_mix_fwd(V[i].data_ptr(), dt, V[i].stride(0), V[i].stride(1), g[i].data_ptr(),
                                 m.data_ptr() if mem else None, C, o[i].data_ptr(), dt, o[i].stride(0), o[i].stride(1),
                                 B, N, n_fft, C, dg, ctypes.c_void_p(st))
        if rc:
            raise RuntimeError(lib.spectre_mix_last_error().decode())

    try:
        for i in range(3):
            run(i)
    except RuntimeError as e:
        return {"error": str(e)}
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        run(i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    es = 4 if dtype == torch.float32 else 2
    nio = min(N, n_fft)
    alg = B * nio * C * es * 2 + B * (C // dg) * (n_fft // 2 + 1) * 8 + ((n_fft // 2 + 1) * C * 8 if mem else 0)
    return {"ms": ms, "GBps": alg / ms / 1e6, "Mtok_s": B * nio / ms / 1e3}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/tune.json")
    args = ap.parse_args()
    lib = _lib.load()
    res = []
    C, dg = 768, 16
    cases = [
        # n_fft, B, tiles to try
        (4096, 64, [0]),
        (1024, 256, [8, 16]),
        (2048, 128, [0]),
        (8192, 32, [0]),
        (16384, 16, [0]),
        (512, 512, [0]),
        (256, 1024, [0]),
        (128, 2048, [0]),
    ]
    for n_fft, B, tiles in cases:
        for tile in tiles:
            for tma, pf in ((1, 1), (1, 0), (0, 0)):
                r = time_case(lib, n_fft, C, dg, B, tile, pf, tma=tma)
                r.update(n_fft=n_fft, B=B, tile=tile, prefetch=pf, tma=tma, dtype="f32")
                print(json.dumps(r), flush=True)
                res.append(r)
    for n_fft, B in [(4096, 64), (1024, 256)]:
        r = time_case(lib, n_fft, C, dg, B, 0, 1, dtype=torch.bfloat16)
        r.update(n_fft=n_fft, B=B, tile=0, prefetch=1, dtype="bf16")
        print(json.dumps(r), flush=True)
        res.append(r)
        r = time_case(lib, n_fft, C, dg, B, 0, 1, mem=True)
        r.update(n_fft=n_fft, B=B, tile=0, prefetch=1, dtype="f32", mem=True)
        print(json.dumps(r), flush=True)
        res.append(r)
    r = time_case(lib, 4096, C, dg, 64, 0, 1, N=3000)
    r.update(n_fft=4096, B=64, N=3000, tile=0, prefetch=1, dtype="f32")
    print(json.dumps(r), flush=True)
    res.append(r)
    lib.spectre_mix_set_tile_channels(0)
    lib.spectre_mix_set_prefetch(1)
    lib.spectre_mix_set_tma(1)
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    json.dump(res, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
