"""Autoregressive decode side: ``PrefixFFTCache`` and ``SpectreHead.decode_step`` on the GPU kernels.

Mirrors ``spectre.py:731-814`` (cache) and ``:562-611`` (head decode).  The running spectrum is updated and the
single output sample is read out in ONE pass over ``prefix_fft`` (``spectre_decode_step``); the gate generator stays
stock PyTorch, exactly as in the forward path.  One cache may hold all heads of a layer (``embed_dim = d``): the gate
row of channel ``c`` is ``c // group_width``.
"""
from __future__ import annotations

import ctypes
import math
from typing import Tuple

import torch

from . import _lib
from .modules import interp_complex_1d
from .ops import rfft_seq


def _stream(dev):
    return ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


class PrefixFFTCache:
    """Sliding-window frequency cache; constructor and attributes as ``spectre.py:745-766``."""

    def __init__(self, n_fft: int, embed_dim: int, device=None):
        if device is None:
            raise ValueError("PrefixFFTCache requires an explicit device parameter. "
                             "Pass device=tensor.device from your input tensors.")
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("fft_b200.PrefixFFTCache is CUDA only (no CPU fallback)")
        self.N, self.d, self.device = n_fft, embed_dim, device
        self.prefix_fft = torch.zeros(n_fft // 2 + 1, embed_dim, dtype=torch.cfloat, device=device)
        self.V_buf = torch.zeros(n_fft, embed_dim, device=device)
        self.Q_buf = torch.zeros_like(self.V_buf)
        self.sum_q = torch.zeros(embed_dim, device=device)
        self.t = -1
        self.freq_k = torch.arange(n_fft // 2 + 1, device=device, dtype=torch.float32)
        self.omega = -2 * math.pi / n_fft
        # per-block partial sums of the read-out (deterministic two-stage reduction); owned by the cache
        self._ws_bytes = _lib.load().spectre_decode_workspace_bytes(n_fft, embed_dim)
        self._ws = torch.empty(self._ws_bytes, dtype=torch.uint8, device=device)

    def prefill(self, Q: torch.Tensor, V: torch.Tensor):
        """Initialise from a prompt of L <= N tokens (``spectre.py:769-783``): spectrum = rfft(pad(V)) on the kernel."""
        L = V.size(0)
        self.prefix_fft.copy_(rfft_seq(V.to(self.device, torch.float32).contiguous(), self.N))
        self.V_buf[:L].copy_(V)
        self.Q_buf[:L].copy_(Q)
        self.sum_q = Q.to(self.device).sum(dim=0)
        self.t = L - 1

    # -- bookkeeping shared by the split and the fused step (ring buffers, running query sum: spectre.py:808-813)
    def _advance(self, q_t: torch.Tensor) -> int:
        """Advance the clock and the query side; returns the ring slot j.  ``V_buf[j]`` still holds the evicted token:
        the kernel reads it from there (no clone) and :meth:`_store_v` overwrites it afterwards, in stream order."""
        self.t += 1
        j = self.t % self.N
        self.Q_buf[j] = q_t
        # spectre.py:810-813 takes `q_old = self.Q_buf[j]` as a VIEW, overwrites the slot, then adds `q_t - q_old`:
        # once the window is full (t >= N) that difference is identically zero, i.e. the running query sum stops
        # moving.  Results must match the reference, so the same happens here.
        if self.t < self.N:
            self.sum_q = self.sum_q + q_t
        return j

    def _store_v(self, j: int, v_t: torch.Tensor):
        self.V_buf[j].copy_(v_t)

    def decode_step(self, q_t: torch.Tensor, v_t: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """Add one token (``spectre.py:786-814``); returns (prefix_fft, sum_q)."""
        v_t = v_t.to(self.device, torch.float32).contiguous()
        j = self._advance(q_t.to(self.device, torch.float32))
        with torch.cuda.device(self.device):
            rc = _lib.load().spectre_decode_update(self.prefix_fft.data_ptr(), v_t.data_ptr(), self.V_buf[j].data_ptr(), self.N,
                                                   self.d, self.t, _stream(self.device))
        _lib.check(rc, "decode_update")
        self._store_v(j, v_t)
        return self.prefix_fft, self.sum_q

    def readout(self, gate_half: torch.Tensor, pos: int) -> torch.Tensor:
        """``pruned_irfft_single(gate_broadcast * prefix_fft, N, pos)`` (``spectre.py:605-609``) on the kernel."""
        gate_half = gate_half.to(self.device, torch.complex64).contiguous()
        out = torch.empty(self.d, device=self.device)
        with torch.cuda.device(self.device):
            rc = _lib.load().spectre_decode_readout(self.prefix_fft.data_ptr(), gate_half.data_ptr(), out.data_ptr(), self.N,
                                                    self.d, self.d // gate_half.shape[0], pos, self._ws.data_ptr(),
                                                    self._ws_bytes, _stream(self.device))
        _lib.check(rc, "decode_readout")
        return out

    def fused_step(self, v_t: torch.Tensor, v_old: torch.Tensor, gate_half: torch.Tensor) -> torch.Tensor:
        """Spectrum update for token ``self.t`` and read-out of sample ``self.t % N`` in one pass over the spectrum."""
        gate_half = gate_half.to(self.device, torch.complex64).contiguous()
        out = torch.empty(self.d, device=self.device)
        with torch.cuda.device(self.device):
            rc = _lib.load().spectre_decode_step(self.prefix_fft.data_ptr(), v_t.data_ptr(), v_old.data_ptr(),
                                                 gate_half.data_ptr(), out.data_ptr(), self.N, self.d,
                                                 self.d // gate_half.shape[0], self.t, self._ws.data_ptr(), self._ws_bytes,
                                                 _stream(self.device))
        _lib.check(rc, "decode_step")
        return out


def decode_gate_torch(head, cache: PrefixFFTCache) -> torch.Tensor:
    """The gate of ``SpectreHead.decode_step`` (``spectre.py:578-598``) in stock PyTorch ops (about twenty small launches);
    kept as the checker of :func:`decode_gate`."""
    descr = head.q_norm((cache.sum_q / cache.N).unsqueeze(0)).squeeze(0)
    gate_rs = head.gate_mlp(descr).view(head.G, head.B, 2)
    gate_anchor = torch.view_as_complex(gate_rs.contiguous())
    gate_half = interp_complex_1d(gate_anchor.unsqueeze(0), size=head.F_half, mode="cubic").squeeze(0)
    gate_half = head.modrelu(gate_half.flatten()).view_as(gate_half)
    k = torch.arange(head.F_half, device=gate_half.device)
    j = cache.t % cache.N
    phase = torch.exp(1j * 2 * math.pi * k * (cache.t - j) / cache.N)
    return gate_half * phase.unsqueeze(0)


def decode_gate(head, cache: PrefixFFTCache) -> torch.Tensor:
    """The gate of ``SpectreHead.decode_step`` (``spectre.py:578-598``).  Running descriptor -> LayerNorm -> gate MLP stay
    stock PyTorch (two GEMVs); cubic interpolation, modReLU and the decode-time positional phase (:586-598) are ONE launch
    (``spectre_decode_gate``).  Returns (G, F_half) complex64."""
    descr = head.q_norm((cache.sum_q / cache.N).unsqueeze(0)).squeeze(0)
    anchors = torch.view_as_complex(head.gate_mlp(descr).float().view(head.G, head.B, 2).contiguous())
    gate = torch.empty(head.G, head.F_half, dtype=torch.complex64, device=anchors.device)
    bias = head.modrelu.bias.detach().float().contiguous()
    eps = head.modrelu.eps.detach().float().reshape(1).expand(head.G).contiguous()
    with torch.cuda.device(anchors.device):
        rc = _lib.load().spectre_decode_gate(anchors.data_ptr(), bias.data_ptr(), eps.data_ptr(), gate.data_ptr(), head.G,
                                             head.G, head.B, head.F_half, cache.t, cache.N, _stream(anchors.device))
    _lib.check(rc, "decode_gate")
    return gate


@torch.no_grad()
def head_decode_step(head, q_t: torch.Tensor, v_t: torch.Tensor, cache: PrefixFFTCache) -> torch.Tensor:
    """``SpectreHead.decode_step`` (``spectre.py:562-611``): one fused kernel pass instead of the phase updates, the
    broadcast multiply and the pruned inverse transform.  Works on our shell and on the reference's ``SpectreHead``."""
    v_t = v_t.to(cache.device, torch.float32).contiguous()
    j = cache._advance(q_t.to(cache.device, torch.float32))               # clock, query ring, running query sum first
    gate_half = decode_gate(head, cache)                                   # depends on sum_q and t only
    out = cache.fused_step(v_t, cache.V_buf[j], gate_half)                 # V_buf[j] = the token being evicted (if any)
    cache._store_v(j, v_t)
    return out
