#!/usr/bin/env python
"""BASELINE config 3: Spectre-base (12 blocks, d=768, 12 heads) seq=4096 bf16 forward on one B200 -- tokens/s."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fft_b200  # noqa: E402

dev = torch.device("cuda")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
model = fft_b200.SpectreBase().to(dev).eval()
tok = torch.randint(0, 32000, (B, 4096), device=dev)
with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
    for _ in range(2):
        model(tok, return_hidden=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        model(tok, return_hidden=True)
    e1.record()
    torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
if len(sys.argv) > 2 and sys.argv[2] == "profile":
    from torch.profiler import ProfilerActivity, profile
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16), profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        model(tok, return_hidden=True)
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=60))
print(f"Spectre-base bf16 fwd: B={B} seq=4096: {ms:.2f} ms/step, {B * 4096 / ms * 1e3:.3e} tokens/s")
