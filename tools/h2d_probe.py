#!/usr/bin/env python
"""Host <-> device copy rates of this box for the e2e leg's buffer sizes (GPU box): plain pinned memory (cudaHostAlloc default)
against write-combined pinned memory for the H2D source, one direction at a time and both at once; host topology facts."""
import ctypes
import json
import os
import subprocess
import time

import torch

rt = ctypes.CDLL("libcudart.so")
N = 1 << 30   # 1 GiB


def host_alloc(nbytes, flags):
    p = ctypes.c_void_p()
    rc = rt.cudaHostAlloc(ctypes.byref(p), ctypes.c_size_t(nbytes), ctypes.c_uint(flags))
    assert rc == 0, rc
    return p


def rate(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return N * reps / (time.perf_counter() - t0) / 1e9


dev = torch.empty(N, dtype=torch.uint8, device="cuda")
dev2 = torch.empty(N, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
out = {}
for name, flags in (("pinned", 0), ("pinned_write_combined", 4)):
    src, dst = host_alloc(N, flags), host_alloc(N, 0)
    ctypes.memset(src, 1, N)

    def h2d():
        rt.cudaMemcpyAsync(ctypes.c_void_p(dev.data_ptr()), src, ctypes.c_size_t(N), 1, ctypes.c_void_p(s1.cuda_stream))

    def d2h():
        rt.cudaMemcpyAsync(dst, ctypes.c_void_p(dev2.data_ptr()), ctypes.c_size_t(N), 2, ctypes.c_void_p(s2.cuda_stream))

    def both():
        h2d()
        d2h()
    out[name] = {"h2d_GBps": round(rate(h2d), 1), "d2h_GBps": round(rate(d2h), 1), "both_each_way_GBps": round(rate(both), 1)}
    rt.cudaFreeHost(src)
    rt.cudaFreeHost(dst)
print(json.dumps(out))
for cmd in (["lscpu"], ["bash", "-c", "ls /sys/devices/system/node/ | head; cat /sys/devices/system/node/node*/cpulist 2>/dev/null"],
            ["bash", "-c", "taskset -p $$; nproc"], ["nvidia-smi", "topo", "-m"]):
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=20).stdout
        print("$", " ".join(cmd))
        keep = [l for l in r.splitlines() if any(k in l for k in ("NUMA", "Socket", "Model name", "CPU(s)", "node", "GPU", "affinity", "-", "NV"))]
        print("\n".join(keep[:40]))
    except Exception as e:  # noqa: BLE001
        print(cmd, "failed", e)
