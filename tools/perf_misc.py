#!/usr/bin/env python
"""Secondary perf figures (GPU box): backward of the mix at the metric shape, decode step latency / bandwidth, the other
BASELINE shapes kernel-only.  Prints one JSON line per measurement."""
import json
import os
import statistics
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fft_b200 import _lib as _lib0  # noqa: E402
if os.environ.get("SPX_ALT"):
    _lib0.LIB_PATH = _lib0.LIB_PATH.replace("libspectre_mix.so", "libspectre_mix_%s.so" % os.environ["SPX_ALT"])
import fft_b200  # noqa: E402
from fft_b200 import ops  # noqa: E402

dev = torch.device("cuda")


def timeit(fn, reps=8, rounds=5):
    fn()
    torch.cuda.synchronize()
    out = []
    for _ in range(rounds):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        out.append(e0.elapsed_time(e1) / reps * 1e3)
        time.sleep(0.2)
    return statistics.median(out), min(out)


def backward():
    B, N, C, dg = int(os.environ.get("PM_BATCH", "64")), 4096, 768, 16
    gen = torch.Generator(device=dev).manual_seed(0)
    V = torch.randn(B, N, C, device=dev, generator=gen)
    dY = torch.randn(B, N, C, device=dev, generator=gen)
    gate = torch.randn(B, C // dg, N // 2 + 1, dtype=torch.cfloat, device=dev, generator=gen)
    unit = B * N * C * 4
    med, best = timeit(lambda: ops._dgate_fused(V, dY, N, dg))
    print(json.dumps({"what": "gate gradient, fused kernel (spectre_mix_dgate)", "B": B, "us": round(med, 1), "us_best": round(best, 1),
                      "algorithmic_GBps": round((2 * unit + gate.numel() * 8) / med / 1e3, 1),
                      "bytes": "V + dY read, dgate written"}), flush=True)

    def two_spectra():
        F_half = N // 2 + 1
        Vf, dYf = ops.rfft_seq(V, N), ops.rfft_seq(dY, N)
        return (torch.conj(Vf) * dYf).view(B, F_half, C // dg, dg).sum(-1).permute(0, 2, 1).contiguous()
    if B <= 64:
        med2, best2 = timeit(two_spectra, reps=3, rounds=3)
        print(json.dumps({"what": "gate gradient, two half spectra + stock reductions (round-1 path)", "B": B, "us": round(med2, 1),
                          "us_best": round(best2, 1)}), flush=True)
    med3, best3 = timeit(lambda: fft_b200.spectral_mix(dY, torch.conj(gate).resolve_conj(), n_fft=N, group_width=dg))
    print(json.dumps({"what": "dV = mix(dY, conj(gate)) incl. the conj copy", "B": B, "us": round(med3, 1), "us_best": round(best3, 1),
                      "algorithmic_GBps": round((2 * unit + gate.numel() * 8) / med3 / 1e3, 1)}), flush=True)
    Vg, gg = V.clone().requires_grad_(), gate.clone().requires_grad_()

    def fwd_bwd():
        Vg.grad = None
        gg.grad = None
        (fft_b200.spectral_mix(Vg, gg, n_fft=N, group_width=dg) * dY).sum().backward()
    med4, best4 = timeit(fwd_bwd, reps=4, rounds=3)
    print(json.dumps({"what": "forward + backward through autograd (includes the elementwise loss)", "B": B, "us": round(med4, 1),
                      "tokens_per_s": round(B * N / med4 * 1e6)}), flush=True)


def decode():
    n, d, G = 4096, 768, 48
    head_d = d
    cache = fft_b200.PrefixFFTCache(n, head_d, device=dev)
    cache.prefill(torch.randn(4000, d, device=dev), torch.randn(4000, d, device=dev))
    gate = torch.randn(G, n // 2 + 1, dtype=torch.cfloat, device=dev)
    q, v = torch.randn(d, device=dev), torch.randn(d, device=dev)

    def step():
        j = cache._advance(q)
        out = cache.fused_step(v, cache.V_buf[j], gate)
        cache._store_v(j, v)
        return out
    med, best = timeit(step, reps=50, rounds=5)
    nbytes = (n // 2 + 1) * d * 8 * 2
    print(json.dumps({"what": "decode step: spectrum update + pruned read-out over prefix_fft (2049 x 768 complex64), all heads of a layer",
                      "us": round(med, 2), "us_best": round(best, 2), "GBps_over_prefix_fft": round(nbytes / med / 1e3, 1),
                      "bytes": "16 B per spectrum element (read + write)"}), flush=True)
    # captured in a CUDA graph (launch overhead out of the picture)
    g = torch.cuda.CUDAGraph()
    vb = cache.V_buf[5]
    out = torch.empty(d, device=dev)
    lib = fft_b200._lib.load() if hasattr(fft_b200, "_lib") else None
    from fft_b200 import _lib
    import ctypes
    lib = _lib.load()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        def raw():
            lib.spectre_decode_step(cache.prefix_fft.data_ptr(), v.data_ptr(), vb.data_ptr(), gate.data_ptr(), out.data_ptr(), n, d,
                                    d // G, 5000, cache._ws.data_ptr(), cache._ws_bytes, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
        raw()
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=s):
            for _ in range(20):
                raw()
    med2, best2 = timeit(lambda: g.replay(), reps=5, rounds=5)
    print(json.dumps({"what": "decode step kernels only (20 steps per CUDA graph replay)", "us_per_step": round(med2 / 20, 2),
                      "GBps_over_prefix_fft": round(nbytes / (med2 / 20) / 1e3, 1)}), flush=True)


def shapes():
    for (B, n, dt) in [(32, 1024, torch.float32), (256, 1024, torch.float32), (128, 2048, torch.float32), (148, 4096, torch.float32),
                       (148, 4096, torch.bfloat16), (32, 8192, torch.float32), (16, 16384, torch.float32)]:
        C, dg = 768, 16
        V = torch.randn(B, n, C, device=dev).to(dt)
        g = torch.randn(B, C // dg, n // 2 + 1, dtype=torch.cfloat, device=dev)
        med, best = timeit(lambda: fft_b200.spectral_mix(V, g, n_fft=n, group_width=dg))
        alg = fft_b200.plan_info(B, n, n, C, dg, dt)["algorithmic_bytes"]
        print(json.dumps({"what": "mix kernel only", "B": B, "n_fft": n, "dtype": str(dt).replace("torch.", ""), "us": round(med, 1),
                          "algorithmic_GBps": round(alg / med / 1e3), "frac_of_6530": round(alg / med / 1e3 / 6530, 3)}), flush=True)
        del V, g


if __name__ == "__main__":
    what = sys.argv[1:] or ["backward", "decode", "shapes"]
    for w in what:
        {"backward": backward, "decode": decode, "shapes": shapes}[w]()
