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
