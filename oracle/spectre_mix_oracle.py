"""CPU oracle for the Spectre spectral-mix forward path.  TEST INFRASTRUCTURE ONLY.

This file restates, on the CPU, the five source lines of the reference that the
CUDA kernel replaces.  It is imported only by ``tests/``, by
``__graft_entry__.smoke()`` and by ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs, always as the checker or the timed CPU baseline and
never from the product path (``fft_b200/``), which fails loudly when the CUDA
library is missing.

Reference lines followed (``/root/reference/spectre.py``):

* ``:506``      ``V_fft = torch.fft.rfft(V, n=n_fft, dim=1)``
* ``:542-543``  ``gate_half.permute(0, 2, 1).repeat_interleave(d_g, dim=-1)``
* ``:545``      ``mixed_half = gate_broadcast * V_fft``
* ``:548-549``  ``mixed_half = mixed_half + memory_fft.unsqueeze(0)``
* ``:551``      ``v_time = torch.fft.irfft(mixed_half, n=n_fft, dim=1)``
* ``:553``      ``v_time[:, :N]``
* ``:703-718``  heads are contiguous channel chunks, looped in Python, ``cat``-ed back

The arithmetic itself lives in a third-party dependency that is not under
``/root/reference``: ``torch.fft`` (ATen ``_fft_r2c`` / ``_fft_c2r`` -> Intel MKL
DFTI on CPU).  The reference pins no version (only ``torch >= 2.2`` at
``spectre.py:22-23``); this image has torch 2.11.0+cu128 / MKL 2024.2.  The
oracle calls the same library through the same call sites, so it is the
reference CPU path, re-assembled outside the ``nn.Module``.

Pinning: the reference ships no tests or golden vectors, so
the oracle is pinned against outputs of the reference itself, captured by
``oracle/make_golden.py`` (which imports ``/root/reference/spectre.py`` in the
build container and hooks ``SpectreHead.forward``) and committed under
``tests/golden/``.  ``tests/test_oracle.py`` checks this file and the
independent C restatement (``spectre_mix_oracle.c``) against those fixtures.
"""
from __future__ import annotations

from typing import Optional

import torch


def mix_one_head(
    V: torch.Tensor,
    gate_half: torch.Tensor,
    n_fft: int,
    d_g: int,
    memory_fft: Optional[torch.Tensor] = None,
) -> torch.Tensor:
    """One head, exactly the op sequence of ``spectre.py:506, :542-553``.

    V          (B, N, d_h) real
    gate_half  (B, G, F_half) complex, G * d_g == d_h
    memory_fft (F_half, d_h) complex or None
    returns    (B, min(N, n_fft), d_h) real
    """
    N = V.shape[1]
    V_fft = torch.fft.rfft(V, n=n_fft, dim=1)                      # :506
    gate_broadcast = gate_half.permute(0, 2, 1)                    # :542
    gate_broadcast = gate_broadcast.repeat_interleave(d_g, dim=-1)  # :543
    mixed_half = gate_broadcast * V_fft                            # :545
    if memory_fft is not None:
        mixed_half = mixed_half + memory_fft.unsqueeze(0)          # :549
    v_time = torch.fft.irfft(mixed_half, n=n_fft, dim=1)           # :551
    return v_time[:, :N]                                           # :553


def mix_head_loop(
    V: torch.Tensor,
    gate: torch.Tensor,
    n_fft: int,
    num_heads: int,
    memory_fft: Optional[torch.Tensor] = None,
) -> torch.Tensor:
    """All heads the way ``SpectreMultiHead.forward`` runs them (``spectre.py:703-718``).

    V     (B, N, C) real, C = num_heads * d_h
    gate  (B, num_heads * G, F_half) complex: head h owns gate rows [h*G, (h+1)*G)
    memory_fft (F_half, C) complex or None, chunked per head like ``:706-707``
    returns (B, min(N, n_fft), C) contiguous (the ``torch.cat`` of ``:718``)
    """
    B, N, C = V.shape
    NG = gate.shape[1]
    assert C % num_heads == 0 and NG % num_heads == 0
    d_h = C // num_heads
    G = NG // num_heads
    assert d_h % G == 0
    d_g = d_h // G
    v_chunks = torch.chunk(V, num_heads, dim=-1)                   # :703 (strided views)
    g_chunks = torch.chunk(gate, num_heads, dim=1)
    m_chunks = (torch.chunk(memory_fft, num_heads, dim=-1)         # :707
                if memory_fft is not None else [None] * num_heads)
    outs = []
    for v, g, m in zip(v_chunks, g_chunks, m_chunks):              # :712-713
        # W_v output is contiguous in the reference (:503); the chunk view here is
        # not, which only changes the copy torch.fft makes, not the values.
        outs.append(mix_one_head(v, g, n_fft, d_g, m))
    return torch.cat(outs, dim=-1)                                 # :718


def mix_flat(
    V: torch.Tensor,
    gate: torch.Tensor,
    n_fft: int,
    group_width: int,
    memory_fft: Optional[torch.Tensor] = None,
    dtype: torch.dtype = torch.float32,
) -> torch.Tensor:
    """Same function without the head loop: channel c uses gate row c // group_width.

    ``dtype=torch.float64`` gives the error-budget oracle (inputs are up-cast,
    the transform runs in double, the result is returned in double).
    """
    cdtype = torch.complex64 if dtype == torch.float32 else torch.complex128
    V = V.to(dtype)
    gate = gate.to(cdtype)
    N = V.shape[1]
    V_fft = torch.fft.rfft(V, n=n_fft, dim=1)
    gb = gate.permute(0, 2, 1).repeat_interleave(group_width, dim=-1)
    mixed = gb * V_fft
    if memory_fft is not None:
        mixed = mixed + memory_fft.to(cdtype).unsqueeze(0)
    return torch.fft.irfft(mixed, n=n_fft, dim=1)[:, :N]


# --------------------------------------------------------------------------
# Decode-side restatements; same "test infrastructure" rule.
# --------------------------------------------------------------------------
def prefill_spectrum(V: torch.Tensor, n_fft: int) -> torch.Tensor:
    """``PrefixFFTCache.prefill`` spectrum, ``spectre.py:774-777``: rfft(pad(V), dim=0)."""
    L = V.shape[0]
    V_pad = torch.nn.functional.pad(V, (0, 0, 0, n_fft - L))
    return torch.fft.rfft(V_pad, dim=0)


def pruned_irfft_single(X_half: torch.Tensor, n: int, pos: int) -> torch.Tensor:
    """One output sample of irfft, following ``spectre.py:614-655`` (even/odd n)."""
    F_half = X_half.shape[0]
    k = torch.arange(F_half, dtype=X_half.real.dtype)
    phase = 2 * torch.pi * k * pos / n
    contrib = X_half.real * torch.cos(phase)[:, None] - X_half.imag * torch.sin(phase)[:, None]
    if n % 2 == 0:
        res = contrib[0] + 2 * contrib[1:-1].sum(0) + contrib[-1] * ((-1) ** pos)
    else:
        res = contrib[0] + 2 * contrib[1:].sum(0)
    return res / n


def cache_update(prefix_fft: torch.Tensor, V_buf: torch.Tensor, v_t: torch.Tensor, t: int, n: int) -> None:
    """In-place spectrum update of ``PrefixFFTCache.decode_step`` for token index t (``spectre.py:795-809``).

    Same complex64 phase arithmetic as the reference: ``exp(1j * omega * freq_k * t)`` with float32 ``freq_k``.
    """
    import math
    omega = -2 * math.pi / n
    freq_k = torch.arange(n // 2 + 1, dtype=torch.float32)
    j = t % n
    if t >= n:
        prefix_fft -= torch.exp(1j * omega * freq_k * j).unsqueeze(-1) * V_buf[j]
    prefix_fft += torch.exp(1j * omega * freq_k * t).unsqueeze(-1) * v_t
    V_buf[j] = v_t


# --------------------------------------------------------------------------
# Gate generator tail; same "test infrastructure" rule.
# --------------------------------------------------------------------------
def gate_tail(anchors: torch.Tensor, bias: torch.Tensor, eps: torch.Tensor,
              pos_phase: Optional[torch.Tensor], F_half: int) -> torch.Tensor:
    """``spectre.py:526-536`` through the same library call sites the reference uses.

    anchors (B, G, Bk) complex; bias (G*F_half,) or (G, F_half); eps scalar tensor or (G,).
    Cubic interpolation = ``F.grid_sample(..., mode="bicubic", padding_mode="border", align_corners=True)`` over a
    height-1 image sampled at ``linspace(-1, 1, F_half)`` (``spectre.py:41-52``), then modReLU (``:109-121``), then the
    positional phase (``:534-536``).  NB ``:41`` reshapes a (B, 2, G, K) stack to (B*G, 2, 1, K): the planes of output
    row j are rows 2j, 2j+1 of [re_0..re_{G-1}, im_0..im_{G-1}], not (re_j, im_j) -- part of the reference's behaviour.
    """
    B, G, K = anchors.shape
    planes = torch.stack([anchors.real, anchors.imag], dim=1).reshape(B * G, 2, 1, K)          # :41
    gx = torch.linspace(-1, 1, F_half).view(1, 1, F_half, 1).expand(B * G, 1, F_half, 1)       # :45-46
    grid = torch.cat([gx, torch.zeros_like(gx)], dim=-1)                                       # :48
    s = torch.nn.functional.grid_sample(planes, grid, mode="bicubic", padding_mode="border", align_corners=True)  # :51
    up = torch.complex(s[:, 0, 0, :], s[:, 1, 0, :]).view(B, G, F_half)                        # :55-60
    mag = torch.abs(up)                                                                        # :110
    e = eps.reshape(-1, 1) if eps.dim() else eps
    scale = torch.relu(mag + bias.reshape(G, F_half)) / torch.sqrt(mag.square() + e.square())  # :115-118
    g = up * scale                                                                             # :121
    if pos_phase is not None:
        g = g * pos_phase.unsqueeze(1 if pos_phase.dim() == 2 else 0)                          # :534-536
    return g


def gate_tail_direct(anchors, bias, eps, pos_phase, F_half: int, dtype="float64"):
    """Independent numpy restatement of the same function from the published definitions (no grid_sample call):
    Keys cubic convolution with A = -0.75, taps floor-1 .. floor+2 clamped to the border, sample positions
    ``k * (Bk - 1) / (F_half - 1)`` (what ``linspace(-1, 1)`` un-normalised with ``align_corners=True`` means)."""
    import numpy as np
    a = np.asarray(anchors).astype(np.complex128 if dtype == "float64" else np.complex64)
    B, G, K = a.shape
    k = np.arange(F_half, dtype=np.float64)
    ix = k * (K - 1) / (F_half - 1)
    fl = np.floor(ix)
    t = ix - fl
    A = -0.75
    c1 = lambda x: ((A + 2) * x - (A + 3)) * x * x + 1          # |x| <= 1
    c2 = lambda x: ((A * x - 5 * A) * x + 8 * A) * x - 4 * A    # 1 < |x| < 2
    w = [c2(t + 1), c1(t), c1(1 - t), c2(2 - t)]
    rows = np.concatenate([a.real, a.imag], axis=1)            # (B, 2G, K): the list R of spectre.py:41's stack
    planes = np.zeros((B, 2 * G, F_half))
    for j in range(4):
        idx = np.clip(fl.astype(np.int64) - 1 + j, 0, K - 1)
        planes += rows[:, :, idx] * w[j]
    up = planes[:, 0::2] + 1j * planes[:, 1::2]                # ... reshaped to (B*G, 2, 1, K): row j = (R[2j], R[2j+1])
    mag = np.abs(up)
    e = np.asarray(eps, dtype=np.float64).reshape(-1, 1) if np.ndim(eps) else float(eps)
    scale = np.maximum(mag + np.asarray(bias, dtype=np.float64).reshape(G, F_half), 0) / np.sqrt(mag ** 2 + e ** 2)
    g = up * scale
    if pos_phase is not None:
        p = np.asarray(pos_phase)
        g = g * (p[:, None, :] if p.ndim == 2 else p)
    return g
