"""Torch-facing operators over the C ABI (``include/spectre_mix.h``).

``spectral_mix`` is what the module shells call where the reference has
``spectre.py:506, :542-553``.  PyTorch is plumbing here -- it owns the device memory
and the stream; the arithmetic is the sm_100a kernel in ``fft_b200/csrc``.
"""
from __future__ import annotations

import contextlib
import ctypes
from typing import Optional

import torch

from . import _lib

_DT = {torch.float32: _lib.F32, torch.bfloat16: _lib.BF16}


def _require_cuda(t: torch.Tensor, name: str):
    if not t.is_cuda:
        raise RuntimeError(
            f"fft_b200: `{name}` is on {t.device}; the spectral-mix path is CUDA (sm_100a) only and has no CPU "
            "fallback. Move the tensors to a B200, or use spectral_mix_host() for host buffers."
        )


def _rows_last_contig(t: torch.Tensor) -> torch.Tensor:
    return t if t.stride(-1) == 1 or t.size(-1) == 1 else t.contiguous()


def _mix_impl(V: torch.Tensor, gate: torch.Tensor, memory: Optional[torch.Tensor], n_fft: int,
              group_width: int) -> torch.Tensor:
    _require_cuda(V, "V")
    if V.dim() != 3:
        raise ValueError(f"V must be (B, N, C), got {tuple(V.shape)}")
    if V.dtype not in _DT:
        raise TypeError(f"V dtype {V.dtype} unsupported (float32 or bfloat16)")
    B, N, C = V.shape
    F_half = n_fft // 2 + 1
    if group_width <= 0 or C % group_width:
        raise ValueError(f"C={C} is not a multiple of group_width={group_width}")
    NG = C // group_width
    if tuple(gate.shape) != (B, NG, F_half):
        raise ValueError(f"gate must be (B, C/group_width, n_fft/2+1) = {(B, NG, F_half)}, got {tuple(gate.shape)}")
    if not gate.is_complex():
        raise TypeError("gate must be complex")
    gate = gate.to(device=V.device, dtype=torch.complex64).contiguous()
    V = _rows_last_contig(V)
    if V.size(-1) == 1 and V.stride(-1) != 1:
        V = V.contiguous()
    mem_ptr, mem_stride = None, 0
    if memory is not None:
        if tuple(memory.shape) != (F_half, C):
            raise ValueError(f"memory must be (n_fft/2+1, C) = {(F_half, C)}, got {tuple(memory.shape)}")
        memory = _rows_last_contig(memory.to(device=V.device, dtype=torch.complex64))
        mem_ptr, mem_stride = memory.data_ptr(), memory.stride(0)
    n_out = min(N, n_fft)
    out = torch.empty((B, n_out, C), dtype=V.dtype, device=V.device)
    if out.numel() == 0:
        return out
    lib = _lib.load()
    with (contextlib.nullcontext() if V.device.index == torch.cuda.current_device() else torch.cuda.device(V.device)):
        # scratch of the long-context path (n_fft > 4096) comes from torch's caching allocator: owned by this call's
        # stream, safe with side streams and under CUDA-graph capture (the kernel never allocates)
        ws_bytes = lib.spectre_mix_workspace_bytes(_DT[V.dtype], B, N, n_fft, C, group_width)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=V.device) if ws_bytes else None
        stream = torch.cuda.current_stream(V.device).cuda_stream
        rc = lib.spectre_mix_fwd_ws(
            V.data_ptr(), _DT[V.dtype], V.stride(0), V.stride(1),
            gate.data_ptr(), mem_ptr, mem_stride,
            out.data_ptr(), _DT[out.dtype], out.stride(0), out.stride(1),
            B, N, n_fft, C, group_width, ws.data_ptr() if ws is not None else None, ws_bytes, ctypes.c_void_p(stream),
        )
    _lib.check(rc, "spectral_mix")
    return out


@torch.library.custom_op("fft_b200::spectral_mix", mutates_args=(), device_types="cuda")
def _spectral_mix_op(V: torch.Tensor, gate: torch.Tensor, memory: Optional[torch.Tensor], n_fft: int,
                     group_width: int) -> torch.Tensor:
    return _mix_impl(V, gate, memory, n_fft, group_width)


@_spectral_mix_op.register_fake
def _(V, gate, memory, n_fft, group_width):
    B, N, C = V.shape
    return V.new_empty((B, min(N, n_fft), C))


def _mix_setup_context(ctx, inputs, output):
    V, gate, memory, n_fft, group_width = inputs
    ctx.n_fft, ctx.group_width = n_fft, group_width
    ctx.N = V.shape[1]
    ctx.has_memory = memory is not None
    ctx.save_for_backward(V, gate)


def _mix_backward(ctx, dY):
    """Adjoint of the mix, built from the same two kernels.

    y = S . irfft . (G * rfft . P) v  with P = zero-pad to n_fft, S = keep N rows.  The gate acts as a real
    circular convolution, so  dV = mix(dY, conj(gate))[:N].  d gate needs the two half spectra:
    dG[b,g,k] = w_k / n * sum_{c in g} conj(Vf[b,k,c]) * dYf[b,k,c]  (w = 1 at DC/Nyquist, 2 elsewhere; the
    imaginary part at DC/Nyquist is dropped like irfft drops it), and dMemory = w_k / n * sum_b dYf.
    """
    V, gate = ctx.saved_tensors
    n_fft, dg, N = ctx.n_fft, ctx.group_width, ctx.N
    dV = dgate = dmem = None
    dY = dY.contiguous()
    if ctx.needs_input_grad[0]:
        # dY has min(N, n_fft) rows; the kernel zero-pads them to n_fft itself (that is S^T)
        full = _mix_impl(dY, torch.conj(gate).resolve_conj(), None, n_fft, dg)
        dV = full[:, :N]
        if N > n_fft:  # rows beyond n_fft never reached the transform (spectre.py:506 truncates)
            dV = torch.nn.functional.pad(dV, (0, 0, 0, N - n_fft))
    need_dmem = ctx.has_memory and ctx.needs_input_grad[2]
    if ctx.needs_input_grad[1] or need_dmem:
        F_half = n_fft // 2 + 1
        w = torch.full((F_half,), 2.0 / n_fft, device=dY.device)
        w[0] = 1.0 / n_fft
        if n_fft % 2 == 0:
            w[-1] = 1.0 / n_fft
        edge = torch.zeros(F_half, dtype=torch.bool, device=dY.device)
        edge[0] = True
        if n_fft % 2 == 0:
            edge[-1] = True
        if need_dmem:
            # the transform is linear: the batch sum of the spectra is the spectrum of the batch sum (one small transform)
            dmem = rfft_seq(dY.float().sum(0, keepdim=True), n_fft)[0] * w[:, None]
            dmem = torch.where(edge[:, None], torch.complex(dmem.real, torch.zeros_like(dmem.real)), dmem)
        if ctx.needs_input_grad[1]:
            dgate = _dgate_fused(V, dY, n_fft, dg)
            if dgate is None:   # layouts without the fused kernel: two half spectra + stock reductions
                dYf = rfft_seq(dY, n_fft) * w[None, :, None]
                Vf = rfft_seq(V, n_fft)
                B, _, C = Vf.shape
                prod = (torch.conj(Vf) * dYf).view(B, F_half, C // dg, dg).sum(-1)   # (B, F, NG)
                prod = torch.where(edge[None, :, None], torch.complex(prod.real, torch.zeros_like(prod.real)), prod)
                dgate = prod.permute(0, 2, 1).contiguous()
    return dV, dgate, dmem, None, None


def _dgate_fused(V: torch.Tensor, dY: torch.Tensor, n_fft: int, group_width: int):
    """``spectre_mix_dgate``: the gate gradient in ONE kernel (both forward transforms, the per-group reduction of
    ``conj(V_fft) * dY_fft`` on chip; no spectrum reaches HBM).  Returns None where no fused variant exists."""
    if V.dtype not in _DT:
        return None
    V = _rows_last_contig(V)
    dY = _rows_last_contig(dY.to(V.dtype))
    B, N, C = V.shape
    dgate = torch.empty((B, C // group_width, n_fft // 2 + 1), dtype=torch.complex64, device=V.device)
    if dgate.numel() == 0:
        return dgate
    lib = _lib.load()
    with torch.cuda.device(V.device):
        stream = torch.cuda.current_stream(V.device).cuda_stream
        rc = lib.spectre_mix_dgate(V.data_ptr(), dY.data_ptr(), _DT[V.dtype], V.stride(0), V.stride(1), dY.stride(0), dY.stride(1),
                                   dgate.data_ptr(), B, N, n_fft, C, group_width, ctypes.c_void_p(stream))
    if rc == 2:          # SPECTRE_MIX_ERR_UNSUPPORTED: no fused variant for this layout
        return None
    _lib.check(rc, "spectral_mix dgate")
    return dgate


_spectral_mix_op.register_autograd(_mix_backward, setup_context=_mix_setup_context)


def spectral_mix(V: torch.Tensor, gate: torch.Tensor, memory: Optional[torch.Tensor] = None, *,
                 n_fft: int, group_width: int) -> torch.Tensor:
    """Fused ``rfft -> gate multiply (+ memory) -> irfft -> [:N]`` over ALL heads in one launch.

    Stands in for ``spectre.py:506, :542-553`` and the head loop of ``:703-718``:

        out[b, n, c] = irfft_n_fft(gate[b, c // group_width, :] * rfft_n_fft(V[b, :, c]) + memory[:, c])[n]

    V       (B, N, C) float32 or bfloat16, CUDA; C = embed_dim (heads are contiguous channel chunks)
    gate    (B, C // group_width, n_fft//2 + 1) complex64; head h owns rows [h*G, (h+1)*G)
    memory  optional (n_fft//2 + 1, C) complex64 (row stride may exceed C)
    returns (B, min(N, n_fft), C) in V's dtype.
    """
    _require_cuda(V, "V")
    needs_grad = torch.is_grad_enabled() and (V.requires_grad or gate.requires_grad or
                                               (memory is not None and memory.requires_grad))
    if not needs_grad and not torch.compiler.is_compiling():
        # inference: straight to the C ABI -- the dispatcher round trip of the custom op is ~10 us, a sixth of the kernel at
        # BASELINE configs[1] (seq 1024, batch 32)
        return _mix_impl(V, gate, memory, int(n_fft), int(group_width))
    return _spectral_mix_op(V, gate, memory, int(n_fft), int(group_width))


def rfft_seq(V: torch.Tensor, n_fft: int) -> torch.Tensor:
    """``torch.fft.rfft(V, n=n_fft, dim=1)`` on the sm_100a kernel (``spectre.py:506`` / ``:777``).

    V (B, N, C) or (N, C); returns complex64 (B, n_fft//2+1, C) (or (n_fft//2+1, C)), contiguous.
    """
    _require_cuda(V, "V")
    squeeze = V.dim() == 2
    if squeeze:
        V = V.unsqueeze(0)
    if V.dtype not in _DT:
        raise TypeError(f"V dtype {V.dtype} unsupported (float32 or bfloat16)")
    V = _rows_last_contig(V)
    B, N, C = V.shape
    spec = torch.empty((B, n_fft // 2 + 1, C), dtype=torch.complex64, device=V.device)
    if spec.numel():
        lib = _lib.load()
        with torch.cuda.device(V.device):
            stream = torch.cuda.current_stream(V.device).cuda_stream
            rc = lib.spectre_rfft_fwd(V.data_ptr(), _DT[V.dtype], V.stride(0), V.stride(1), spec.data_ptr(),
                                      B, N, n_fft, C, ctypes.c_void_p(stream))
        _lib.check(rc, "rfft_seq")
    return spec[0] if squeeze else spec


# ----------------------------------------------------------------------------- gate generator tail
def _gate_args(anchors, bias, eps, pos_phase, F_half: int, G: int):
    """Validate / normalise the gate generator's inputs; returns (anchors, bias, eps, pos_phase or None, pos_stride_b)."""
    _require_cuda(anchors, "anchors")
    if anchors.dim() != 3 or not anchors.is_complex():
        raise ValueError(f"anchors must be complex (B, NG, Bk), got {tuple(anchors.shape)} {anchors.dtype}")
    B, NG, Bk = anchors.shape
    if G <= 0 or NG % G:
        raise ValueError(f"NG={NG} is not a multiple of the gate rows per head G={G}")
    if tuple(bias.shape) != (NG, F_half):
        raise ValueError(f"bias must be (NG, F_half) = {(NG, F_half)}, got {tuple(bias.shape)}")
    if tuple(eps.shape) != (NG,):
        raise ValueError(f"eps must be (NG,) = {(NG,)}, got {tuple(eps.shape)}")
    anchors = anchors.to(torch.complex64).contiguous()
    bias = bias.to(device=anchors.device, dtype=torch.float32).contiguous()
    eps = eps.to(device=anchors.device, dtype=torch.float32).contiguous()
    pos_stride = 0
    if pos_phase is not None:
        pos_phase = pos_phase.to(device=anchors.device, dtype=torch.complex64)
        if pos_phase.dim() == 1:
            pos_phase = pos_phase.unsqueeze(0)
        if pos_phase.dim() != 2 or pos_phase.shape[-1] != F_half or pos_phase.shape[0] not in (1, B):
            raise ValueError(f"pos_phase must be (F_half,), (1, F_half) or (B, F_half), got {tuple(pos_phase.shape)}")
        pos_phase = pos_phase.contiguous()
        pos_stride = F_half if pos_phase.shape[0] == B and B > 1 else 0
    return anchors, bias, eps, pos_phase, pos_stride


def _gate_expand_impl(anchors: torch.Tensor, bias: torch.Tensor, eps: torch.Tensor, pos_phase: Optional[torch.Tensor],
                      F_half: int, G: int) -> torch.Tensor:
    anchors, bias, eps, pos_phase, pos_stride = _gate_args(anchors, bias, eps, pos_phase, F_half, G)
    B, NG, Bk = anchors.shape
    pos_ptr = None if pos_phase is None else pos_phase.data_ptr()
    gate = torch.empty((B, NG, F_half), dtype=torch.complex64, device=anchors.device)
    if gate.numel():
        lib = _lib.load()
        with torch.cuda.device(anchors.device):
            stream = torch.cuda.current_stream(anchors.device).cuda_stream
            rc = lib.spectre_gate_expand(anchors.data_ptr(), bias.data_ptr(), eps.data_ptr(), pos_ptr, pos_stride,
                                         gate.data_ptr(), B, NG, G, Bk, F_half, ctypes.c_void_p(stream))
        _lib.check(rc, "gate_expand")
    return gate


def _gate_expand_torch(anchors, bias, eps, pos_phase, F_half, G):
    """The same function in stock PyTorch ops (spectre.py:526-536); used to differentiate the fused kernel."""
    from .modules import interp_complex_1d
    B, NG, Bk = anchors.shape
    # per head, as the reference calls it (the real/imag row shuffle of spectre.py:41 stays inside a head)
    g = interp_complex_1d(anchors.reshape(B * (NG // G), G, Bk), size=F_half, mode="cubic").reshape(B, NG, F_half)
    mag = torch.abs(g)
    scale = torch.relu(mag + bias) / torch.sqrt(mag.square() + eps[:, None].square())
    g = g * scale
    if pos_phase is not None:
        g = g * (pos_phase.unsqueeze(1) if pos_phase.dim() == 2 else pos_phase)
    return g


class _GateExpand(torch.autograd.Function):
    @staticmethod
    def forward(ctx, anchors, bias, eps, pos_phase, F_half, G):
        ctx.F_half, ctx.G = F_half, G
        ctx.save_for_backward(anchors, bias, eps, pos_phase)
        return _gate_expand_impl(anchors, bias, eps, pos_phase, F_half, G)

    @staticmethod
    def backward(ctx, dgate):
        anchors, bias, eps, pos_phase = ctx.saved_tensors
        with torch.enable_grad():
            a = anchors.detach().requires_grad_(ctx.needs_input_grad[0])
            b = bias.detach().requires_grad_(ctx.needs_input_grad[1])
            g = _gate_expand_torch(a, b, eps, pos_phase, ctx.F_half, ctx.G)
            wanted = [t for t, need in ((a, ctx.needs_input_grad[0]), (b, ctx.needs_input_grad[1])) if need]
            grads = list(torch.autograd.grad(g, wanted, dgate)) if wanted else []
        da = grads.pop(0) if ctx.needs_input_grad[0] else None
        db = grads.pop(0) if ctx.needs_input_grad[1] else None
        return da, db, None, None, None, None


def gate_expand(anchors: torch.Tensor, bias: torch.Tensor, eps: torch.Tensor, pos_phase: Optional[torch.Tensor] = None,
                *, F_half: int, G: int) -> torch.Tensor:
    """Gate generator tail for ALL heads in one launch: cubic interpolation of the anchors to ``F_half`` bins,
    modReLU, optional positional phase (``spectre.py:526-536``, ``:26-61``, ``:109-121``).

    anchors   (B, NG, Bk) complex64 -- ``gate_mlp`` output viewed as complex, heads stacked along NG
    bias      (NG, F_half) float32  -- ``modrelu.bias`` of every head, stacked
    eps       (NG,) float32         -- ``modrelu.eps`` of the head each row belongs to
    pos_phase optional complex (F_half,), (1, F_half) or (B, F_half)
    G         gate rows per head: the reference interpolates each head's (G, Bk) block through a reshape that pairs
              consecutive rows of [re_0..re_{G-1}, im_0..im_{G-1}] as (real, imag) planes (spectre.py:41); kept as is
    returns   (B, NG, F_half) complex64.  Differentiable w.r.t. anchors and bias (backward in stock PyTorch ops).
    """
    _require_cuda(anchors, "anchors")
    return _GateExpand.apply(anchors, bias, eps, pos_phase, int(F_half), int(G))


# ----------------------------------------------------------------------------- mix with the gate generator fused in (8f-2)
def _mix_anchors_impl(V, anchors, bias, eps, pos_phase, memory, n_fft: int, group_width: int, G: int) -> torch.Tensor:
    _require_cuda(V, "V")
    if V.dim() != 3:
        raise ValueError(f"V must be (B, N, C), got {tuple(V.shape)}")
    if V.dtype not in _DT:
        raise TypeError(f"V dtype {V.dtype} unsupported (float32 or bfloat16)")
    B, N, C = V.shape
    F_half = n_fft // 2 + 1
    if group_width <= 0 or C % group_width:
        raise ValueError(f"C={C} is not a multiple of group_width={group_width}")
    NG = C // group_width
    anchors, bias, eps, pos_phase, pos_stride = _gate_args(anchors, bias, eps, pos_phase, F_half, G)
    if tuple(anchors.shape[:2]) != (B, NG):
        raise ValueError(f"anchors must be (B, C/group_width, Bk) = ({B}, {NG}, Bk), got {tuple(anchors.shape)}")
    V = _rows_last_contig(V)
    es = V.element_size()
    if group_width % 4 == 0 and (V.data_ptr() % (4 * es) or V.stride(0) % 4 or V.stride(1) % 4):
        V = V.contiguous()      # keep the packed layout (the one with the fused gate generator) reachable
    mem_ptr, mem_stride = None, 0
    if memory is not None:
        if tuple(memory.shape) != (F_half, C):
            raise ValueError(f"memory must be (n_fft/2+1, C) = {(F_half, C)}, got {tuple(memory.shape)}")
        memory = _rows_last_contig(memory.to(device=V.device, dtype=torch.complex64))
        mem_ptr, mem_stride = memory.data_ptr(), memory.stride(0)
    out = torch.empty((B, min(N, n_fft), C), dtype=V.dtype, device=V.device)
    if out.numel() == 0:
        return out
    lib = _lib.load()
    with torch.cuda.device(V.device):
        ws_bytes = lib.spectre_mix_anchors_workspace_bytes(_DT[V.dtype], B, N, n_fft, C, group_width)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=V.device) if ws_bytes else None
        stream = torch.cuda.current_stream(V.device).cuda_stream
        rc = lib.spectre_mix_fwd_anchors(
            V.data_ptr(), _DT[V.dtype], V.stride(0), V.stride(1),
            anchors.data_ptr(), bias.data_ptr(), eps.data_ptr(), None if pos_phase is None else pos_phase.data_ptr(), pos_stride,
            G, anchors.shape[2], mem_ptr, mem_stride,
            out.data_ptr(), _DT[out.dtype], out.stride(0), out.stride(1),
            B, N, n_fft, C, group_width, ws.data_ptr() if ws is not None else None, ws_bytes, ctypes.c_void_p(stream),
        )
    _lib.check(rc, "spectral_mix_anchors")
    return out


@torch.library.custom_op("fft_b200::spectral_mix_anchors", mutates_args=(), device_types="cuda")
def _spectral_mix_anchors_op(V: torch.Tensor, anchors: torch.Tensor, bias: torch.Tensor, eps: torch.Tensor,
                             pos_phase: Optional[torch.Tensor], memory: Optional[torch.Tensor], n_fft: int, group_width: int,
                             G: int) -> torch.Tensor:
    return _mix_anchors_impl(V, anchors, bias, eps, pos_phase, memory, n_fft, group_width, G)


@_spectral_mix_anchors_op.register_fake
def _(V, anchors, bias, eps, pos_phase, memory, n_fft, group_width, G):
    B, N, C = V.shape
    return V.new_empty((B, min(N, n_fft), C))


def _mix_anchors_setup_context(ctx, inputs, output):
    V, anchors, bias, eps, pos_phase, memory, n_fft, group_width, G = inputs
    ctx.n_fft, ctx.group_width, ctx.G = n_fft, group_width, G
    ctx.has_memory = memory is not None
    ctx.save_for_backward(V, anchors, bias, eps, pos_phase)


def _mix_anchors_backward(ctx, dY):
    """Backward of the fused op = backward of ``spectral_mix`` chained with the gate generator's: the gate is materialised
    here (backward only), d gate comes from the mix adjoint, and anchors / bias receive it through the stock-op formula."""
    V, anchors, bias, eps, pos_phase = ctx.saved_tensors
    F_half = ctx.n_fft // 2 + 1
    need_gate_grads = ctx.needs_input_grad[1] or ctx.needs_input_grad[2]
    with torch.enable_grad():
        a = anchors.detach().requires_grad_(ctx.needs_input_grad[1])
        b = bias.detach().requires_grad_(ctx.needs_input_grad[2])
        if need_gate_grads:
            gate = _gate_expand_torch(a, b, eps, pos_phase, F_half, ctx.G)
        else:
            gate = _gate_expand_impl(anchors, bias, eps, pos_phase, F_half, ctx.G)
    inner = type("Ctx", (), {})()
    inner.n_fft, inner.group_width, inner.N, inner.has_memory = ctx.n_fft, ctx.group_width, V.shape[1], ctx.has_memory
    inner.saved_tensors = (V, gate.detach())
    inner.needs_input_grad = (ctx.needs_input_grad[0], need_gate_grads, ctx.needs_input_grad[5], False, False)
    dV, dgate, dmem, _, _ = _mix_backward(inner, dY)
    da = db = None
    if need_gate_grads:
        wanted = [t for t, need in ((a, ctx.needs_input_grad[1]), (b, ctx.needs_input_grad[2])) if need]
        grads = list(torch.autograd.grad(gate, wanted, dgate))
        da = grads.pop(0) if ctx.needs_input_grad[1] else None
        db = grads.pop(0) if ctx.needs_input_grad[2] else None
    return dV, da, db, None, None, dmem, None, None, None


_spectral_mix_anchors_op.register_autograd(_mix_anchors_backward, setup_context=_mix_anchors_setup_context)


def spectral_mix_anchors(V: torch.Tensor, anchors: torch.Tensor, bias: torch.Tensor, eps: torch.Tensor,
                         pos_phase: Optional[torch.Tensor] = None, memory: Optional[torch.Tensor] = None, *,
                         n_fft: int, group_width: int, G: int) -> torch.Tensor:
    """``spectral_mix(V, gate_expand(anchors, bias, eps, pos_phase), memory)`` in ONE launch: the gate generator's tail
    (``spectre.py:526-536``) is evaluated inside the mix kernel's gate staging and the ``(B, NG, F_half)`` gate is never
    materialised.  Arguments as :func:`gate_expand` and :func:`spectral_mix`."""
    _require_cuda(V, "V")
    needs_grad = torch.is_grad_enabled() and any(
        t is not None and t.requires_grad for t in (V, anchors, bias, eps, pos_phase, memory))
    if not needs_grad and not torch.compiler.is_compiling():   # inference: skip the dispatcher (see spectral_mix)
        return _mix_anchors_impl(V, anchors, bias, eps, pos_phase, memory, int(n_fft), int(group_width), int(G))
    return _spectral_mix_anchors_op(V, anchors, bias, eps, pos_phase, memory, int(n_fft), int(group_width), int(G))


def spectral_mix_host(V: torch.Tensor, gate: torch.Tensor, memory: Optional[torch.Tensor] = None, *,
                      n_fft: int, group_width: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Same function on HOST tensors through ``spectre_mix_fwd_host`` (copies inside the call).

    This is the entry a non-CUDA caller binds and the one ``bench.py`` times for ``e2e``.  float32 only;
    pinned tensors copy at full PCIe rate.
    """
    if V.is_cuda or gate.is_cuda:
        raise RuntimeError("spectral_mix_host takes CPU tensors; use spectral_mix for CUDA tensors")
    if V.dim() != 3:
        raise ValueError(f"V must be (B, N, C), got {tuple(V.shape)}")
    V = V.contiguous().float()
    gate = gate.contiguous().to(torch.complex64)
    B, N, C = V.shape
    n_out = min(N, n_fft)
    F_half = n_fft // 2 + 1
    if group_width <= 0 or C % group_width:
        raise ValueError(f"C={C} is not a multiple of group_width={group_width}")
    if tuple(gate.shape) != (B, C // group_width, F_half):
        raise ValueError(f"gate must be {(B, C // group_width, F_half)}, got {tuple(gate.shape)}")
    if out is None:
        out = torch.empty((B, n_out, C), dtype=torch.float32)
    elif (out.is_cuda or out.dtype != torch.float32 or tuple(out.shape) != (B, n_out, C) or not out.is_contiguous()):
        # the C entry writes B * n_out * C floats through this pointer: anything else would be an out-of-bounds write
        raise ValueError(f"out must be a contiguous CPU float32 tensor of shape {(B, n_out, C)}, got "
                         f"{tuple(out.shape)} {out.dtype} on {out.device} (contiguous={out.is_contiguous()})")
    mem_ptr = None
    if memory is not None:
        if tuple(memory.shape) != (F_half, C):
            raise ValueError(f"memory must be {(F_half, C)}, got {tuple(memory.shape)}")
        memory = memory.contiguous().to(torch.complex64)
        mem_ptr = memory.data_ptr()
    lib = _lib.load()
    rc = lib.spectre_mix_fwd_host(V.data_ptr(), gate.data_ptr(), mem_ptr, out.data_ptr(), B, N, n_fft, C, group_width)
    _lib.check(rc, "spectral_mix_host")
    return out


def plan_info(B: int, N: int, n_fft: int, C: int, group_width: int, dtype: torch.dtype = torch.float32,
              has_memory: bool = False) -> dict:
    """How the library will run a problem (tile shape, grid, shared memory, algorithmic bytes)."""
    info = _lib.PlanInfo()
    rc = _lib.load().spectre_mix_plan(_DT[dtype], _DT[dtype], int(has_memory), B, N, n_fft, C, group_width,
                                      ctypes.byref(info))
    _lib.check(rc, "plan_info")
    return {
        "n_fft": info.n_fft, "radix": [r for r in info.radix if r > 1], "tile_channels": info.tile_channels,
        "threads": info.threads, "ctas_per_sm": info.ctas_per_sm, "smem_bytes": info.smem_bytes,
        "grid": info.grid, "launches": info.launches, "algorithmic_bytes": info.algorithmic_bytes,
        "workspace_bytes": info.workspace_bytes, "dit": info.dit,
    }
