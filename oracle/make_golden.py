#!/usr/bin/env python
"""Generate tests/golden/*.npz from the UNMODIFIED reference (build container only).

Run here, where /root/reference exists:  ``python oracle/make_golden.py``.
It imports ``/root/reference/spectre.py``, runs the stock modules on seeded CPU
inputs and records, through forward hooks on ``SpectreHead`` (no code of the
reference is changed or copied), the tensors that cross the boundary the CUDA
kernel replaces:

    V          output of ``head.W_v``             (spectre.py:503)
    gate_half  output of ``head.modrelu`` reshaped (spectre.py:531)
    memory     the ``memory_fft=`` chunk passed to the head (spectre.py:706-713)
    out        the head's return value            (spectre.py:553)

plus block-level fixtures (state_dict, input, output) for the drop-in module
test.  The GPU box has no /root/reference, so the vectors are committed.
"""
from __future__ import annotations

import math
import os
import sys

import numpy as np
import torch

REF = os.environ.get("SPECTRE_REFERENCE_DIR", "/root/reference")
sys.path.insert(0, REF)
import spectre as ref  # noqa: E402  (the reference, unmodified)

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def capture_heads(mh, x, memory_fft=None):
    """Run a stock SpectreMultiHead and capture per-head boundary tensors."""
    rec = [dict() for _ in mh.heads]
    hooks = []
    for i, h in enumerate(mh.heads):
        hooks.append(h.W_v.register_forward_hook(
            lambda m, a, o, i=i: rec[i].__setitem__("V", o.detach().clone())))
        hooks.append(h.modrelu.register_forward_hook(
            lambda m, a, o, i=i, h=h: rec[i].__setitem__(
                "gate", o.detach().clone().view(o.shape[0], h.G, h.F_half))))
        hooks.append(h.register_forward_pre_hook(
            lambda m, a, kw, i=i: rec[i].__setitem__(
                "mem", None if kw.get("memory_fft") is None else kw["memory_fft"].detach().clone()),
            with_kwargs=True))
        hooks.append(h.register_forward_hook(
            lambda m, a, o, i=i: rec[i].__setitem__("out", o[0].detach().clone())))
    with torch.no_grad():
        y = mh(x, memory_fft=memory_fft)
    for hk in hooks:
        hk.remove()
    return rec, y


def flat_case(rec):
    """Concatenate per-head captures into the all-heads form the kernel takes."""
    V = torch.cat([r["V"] for r in rec], dim=-1)
    gate = torch.cat([r["gate"] for r in rec], dim=1)
    out = torch.cat([r["out"] for r in rec], dim=-1)
    mem = None if rec[0]["mem"] is None else torch.cat([r["mem"] for r in rec], dim=-1)
    return V, gate, mem, out


def save_mix_case(name, rec, n_fft, num_heads, group_width):
    V, gate, mem, out = flat_case(rec)
    d = dict(V=V.numpy(), gate=gate.numpy(), out=out.numpy(),
             n_fft=np.int64(n_fft), num_heads=np.int64(num_heads), group_width=np.int64(group_width))
    if mem is not None:
        d["mem"] = mem.numpy()
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **d)
    print(f"{name}: V{tuple(V.shape)} gate{tuple(gate.shape)} mem={None if mem is None else tuple(mem.shape)} "
          f"out{tuple(out.shape)} -> {os.path.getsize(path)/1024:.0f} KiB")


def block_mem(block):
    """Memory exactly as SpectreBlock.forward prepares it (spectre.py:973-977)."""
    m = block.memory_fft
    if m is not None and block.memory_freq_bins < block.full_freq_bins:
        m = torch.nn.functional.pad(m, (0, 0, 0, block.full_freq_bins - block.memory_freq_bins))
    return m


def main():
    os.makedirs(OUT, exist_ok=True)
    kw = dict(pooling_type="mean", wavelet_on_rate=0.0, use_toeplitz=False)

    # (1) BASELINE config 1 shape: B=2, seq=128, d=64, 4 heads, G=4 (d_g=4), full-size memory
    torch.manual_seed(0)
    blk = ref.SpectreBlock(64, 4, 128, memory_size=1, **kw).eval()
    x = torch.randn(2, 128, 64)
    rec, _ = capture_heads(blk.mix, blk.ln1(x), block_mem(blk))
    save_mix_case("mix_b2_n128_c64_mem", rec, 128, 4, 4)

    # (2) no memory, N < n_fft (zero padding, linear convolution; output has N rows)
    torch.manual_seed(1)
    mh = ref.SpectreMultiHead(64, 4, 128, **kw).eval()
    rec, _ = capture_heads(mh, torch.randn(2, 100, 64))
    save_mix_case("mix_b2_n100_nfft128_c64", rec, 128, 4, 4)

    # (3) truncated memory (memory_size=17 < F_half=65 -> zero-padded high bins), d_g = 8
    torch.manual_seed(2)
    blk3 = ref.SpectreBlock(64, 2, 128, memory_size=17, num_groups=4, **kw).eval()
    x3 = torch.randn(3, 128, 64)
    rec, _ = capture_heads(blk3.mix, blk3.ln1(x3), block_mem(blk3))
    save_mix_case("mix_b3_n128_c64_memtrunc", rec, 128, 2, 8)

    # (4) longer transform, three radix stages in the kernel: n_fft=1024, d=64, 2 heads, d_g=8
    torch.manual_seed(3)
    mh4 = ref.SpectreMultiHead(64, 2, 1024, **kw).eval()
    rec, _ = capture_heads(mh4, torch.randn(1, 1024, 64))
    save_mix_case("mix_b1_n1024_c64", rec, 1024, 2, 8)

    # (5) the metric's transform length, narrow: n_fft=4096, d=32, 2 heads (d_h=16, d_g=4), N=3000 < n_fft
    torch.manual_seed(4)
    mh5 = ref.SpectreMultiHead(32, 2, 4096, **kw).eval()
    rec, _ = capture_heads(mh5, torch.randn(1, 3000, 32))
    save_mix_case("mix_b1_n3000_nfft4096_c32", rec, 4096, 2, 4)

    # (6) group width that is even but not a multiple of 4 (d_h=24, G=4 -> d_g=6), and odd (d_h=12 -> d_g=3)
    torch.manual_seed(5)
    mh6 = ref.SpectreMultiHead(48, 2, 64, **kw).eval()
    rec, _ = capture_heads(mh6, torch.randn(2, 64, 48))
    save_mix_case("mix_b2_n64_c48_dg6", rec, 64, 2, 6)
    torch.manual_seed(6)
    mh7 = ref.SpectreMultiHead(24, 2, 64, **kw).eval()
    rec, _ = capture_heads(mh7, torch.randn(2, 64, 24))
    save_mix_case("mix_b2_n64_c24_dg3", rec, 64, 2, 3)

    # (7) block-level drop-in fixture: state_dict + input + output of the stock SpectreBlock
    for name, seed, memsz in (("block_d64_h4_n128_mem", 10, 1), ("block_d64_h4_n128", 11, 0)):
        torch.manual_seed(seed)
        b = ref.SpectreBlock(64, 4, 128, memory_size=memsz, **kw).eval()
        xin = torch.randn(2, 128, 64)
        with torch.no_grad():
            yout = b(xin)
        d = {"sd::" + k: v.detach().numpy() for k, v in b.state_dict().items()}
        d["x"] = xin.numpy()
        d["y"] = yout.numpy()
        d["memory_size"] = np.int64(memsz)
        path = os.path.join(OUT, name + ".npz")
        np.savez_compressed(path, **d)
        print(f"{name}: {len(b.state_dict())} tensors -> {os.path.getsize(path)/1024:.0f} KiB")

    # (8) decode-side fixtures: prefill spectrum and pruned irfft
    torch.manual_seed(20)
    cache = ref.PrefixFFTCache(256, 32, device=torch.device("cpu"))
    Q = torch.randn(200, 32)
    Vp = torch.randn(200, 32)
    cache.prefill(Q, Vp)
    Xh = torch.randn(129, 32, dtype=torch.cfloat)
    pr = torch.stack([ref.pruned_irfft_single(Xh, 256, p) for p in (0, 1, 7, 255)])
    np.savez_compressed(os.path.join(OUT, "decode_prefill_n256_d32.npz"),
                        V=Vp.numpy(), prefix_fft=cache.prefix_fft.numpy(),
                        X_half=Xh.numpy(), pruned_pos=np.array([0, 1, 7, 255]), pruned=pr.numpy())
    print("decode_prefill_n256_d32 written")

    # (9) decode sequence through the stock SpectreHead.decode_step + PrefixFFTCache (spectre.py:562-611, :786-814):
    # prompt of 200 tokens, then 100 single-token steps (crosses t >= N = 256, so eviction is exercised)
    torch.manual_seed(21)
    head = ref.SpectreHead(32, 256, pooling_type="mean").eval()
    cache = ref.PrefixFFTCache(256, 32, device=torch.device("cpu"))
    with torch.no_grad():
        xp = torch.randn(200, 32)
        Qp, Vp = head.W_q(xp), head.W_v(xp)
        cache.prefill(Qp, Vp)
        xs = torch.randn(100, 32)
        qs, vs = head.W_q(xs), head.W_v(xs)
        outs = torch.stack([head.decode_step(qs[i], vs[i], cache) for i in range(100)])
    d = {"sd::" + k: v.detach().numpy() for k, v in head.state_dict().items()}
    d.update(Qp=Qp.numpy(), Vp=Vp.numpy(), qs=qs.numpy(), vs=vs.numpy(), outs=outs.numpy(),
             prefix_fft=cache.prefix_fft.numpy(), sum_q=cache.sum_q.numpy(), t=np.int64(cache.t))
    np.savez_compressed(os.path.join(OUT, "decode_seq_n256_d32.npz"), **d)
    print("decode_seq_n256_d32 written", outs.shape)

    # (10) gate generator tail: the reference's own interp_complex_1d (spectre.py:26-61) and ComplexModReLU
    # (:109-121) on random anchors, with a bias spread that exercises the ReLU cut-off, and the positional phase of :534-536
    for name, seed, n_fft, G, Bsz in (("gate_n128_g4", 30, 128, 4, 3), ("gate_n4096_g4", 31, 4096, 4, 2),
                                      ("gate_n1000_g3", 32, 1000, 3, 2)):
        torch.manual_seed(seed)
        F_half = n_fft // 2 + 1
        Bk = max(4, int(math.sqrt(F_half)))
        anchors = torch.randn(Bsz, G, Bk, dtype=torch.cfloat)
        mr = ref.ComplexModReLU(F_half * G)
        with torch.no_grad():
            mr.bias.copy_(-0.1 + 0.6 * torch.randn(F_half * G))
            up = ref.interp_complex_1d(anchors, size=F_half, mode="cubic")
            gate = mr(up.reshape(Bsz, -1)).view_as(up)
            ang1 = torch.rand(1, F_half) * 6.28
            angB = torch.rand(Bsz, F_half) * 6.28
            pos1, posB = torch.polar(torch.ones_like(ang1), ang1), torch.polar(torch.ones_like(angB), angB)
            gate_pos1 = gate * pos1.unsqueeze(1)        # spectre.py:536 (pos_phase.dim() == 2)
            gate_posB = gate * posB.unsqueeze(1)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), anchors=anchors.numpy(), bias=mr.bias.detach().numpy(),
                            eps=mr.eps.numpy(), interp=up.numpy(), gate=gate.numpy(), pos1=pos1.numpy(), posB=posB.numpy(),
                            gate_pos1=gate_pos1.numpy(), gate_posB=gate_posB.numpy(), n_fft=np.int64(n_fft))
        print(f"{name}: anchors{tuple(anchors.shape)} -> gate{tuple(gate.shape)}")


if __name__ == "__main__":
    main()
