"""GPU parity tests: the sm_100a kernel, called through the C ABI, against the CPU oracle.

Tolerances (SURVEY 8c / BASELINE north_star): fp32 rel-L2 <= 1e-5 and max-abs <= 1e-4 * max|y| (the
north_star's outer bound is 1e-3 relative); bf16 I/O rel-L2 <= 1e-2.  The oracle's own fp32-vs-fp64
error at n_fft=4096 is rel-L2 1.9e-7.
"""
import numpy as np
import pytest
import torch

from conftest import golden_names, load_golden, max_abs_rel, rel_l2

pytestmark = pytest.mark.gpu

REL_L2_F32 = 1e-5
MAX_ABS_F32 = 1e-4
REL_L2_BF16 = 1e-2


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def fb():
    import fft_b200
    from fft_b200 import _lib
    _lib.load()  # fail loudly if the CUDA library is missing
    return fft_b200


@pytest.fixture(scope="module")
def oracle():
    from oracle import spectre_mix_oracle
    return spectre_mix_oracle


def _check(got, want, rl2=REL_L2_F32, mabs=MAX_ABS_F32):
    got = got.detach().float().cpu().numpy()
    want = np.asarray(want, dtype=np.float32)
    assert got.shape == want.shape, (got.shape, want.shape)
    assert np.isfinite(got).all()
    e1, e2 = rel_l2(got, want), max_abs_rel(got, want)
    assert e1 <= rl2 and e2 <= mabs, f"rel-L2 {e1:.3e} (<= {rl2}), max-abs/max {e2:.3e} (<= {mabs})"
    return e1


# --------------------------------------------------------------------------- reference-generated goldens
@pytest.mark.parametrize("name", golden_names("mix_"))
def test_goldens_from_reference(name, fb, dev):
    g = load_golden(name)
    mem = None if "mem" not in g else torch.from_numpy(g["mem"]).to(dev)
    y = fb.spectral_mix(torch.from_numpy(g["V"]).to(dev), torch.from_numpy(g["gate"]).to(dev), mem,
                        n_fft=int(g["n_fft"]), group_width=int(g["group_width"]))
    _check(y, g["out"])


# --------------------------------------------------------------------------- seeded sweeps against the oracle
def _rand_case(B, N, n_fft, C, dg, with_mem, seed):
    gen = torch.Generator().manual_seed(seed)
    V = torch.randn(B, N, C, generator=gen)
    gate = torch.randn(B, C // dg, n_fft // 2 + 1, dtype=torch.cfloat, generator=gen)
    mem = torch.randn(n_fft // 2 + 1, C, dtype=torch.cfloat, generator=gen) / np.sqrt(C) if with_mem else None
    return V, gate, mem


SWEEP = [
    # B, N, n_fft, C, d_g, mem
    (2, 32, 32, 16, 4, False),
    (2, 64, 64, 16, 8, True),
    (3, 128, 128, 64, 4, True),          # BASELINE config 1 shape (per block: d=64, 4 heads, d_g=4)
    (2, 256, 256, 32, 16, False),
    (2, 512, 512, 48, 16, True),
    (2, 1024, 1024, 64, 16, True),
    (2, 2048, 2048, 32, 16, False),
    (2, 4096, 4096, 32, 16, True),
    (1, 8192, 8192, 16, 16, False),
    (1, 16384, 16384, 16, 16, True),     # BASELINE config 4 transform length
    # ragged: zero padding (N < n_fft), truncation (N > n_fft), one row
    (2, 3000, 4096, 16, 16, False),
    (2, 1, 1024, 16, 16, True),
    (2, 700, 512, 16, 16, False),
    (1, 5000, 4096, 32, 16, True),
    # channel counts that leave a partial tile, group widths for every element mode
    (2, 1024, 1024, 24, 8, True),        # QUAD, C not a multiple of the tile
    (2, 4096, 4096, 40, 8, False),
    (3, 4096, 4096, 96, 12, True),       # QUAD with a group width that is not a power of two (division path of the tile loop)
    (2, 256, 256, 36, 6, True),          # PAIR (d_g even, not /4)
    (2, 1024, 1024, 20, 10, False),
    (2, 256, 256, 15, 3, True),          # REAL (odd group width)
    (2, 2048, 2048, 7, 1, False),        # one gate per channel
    (1, 4096, 4096, 768, 16, False),     # Spectre-base row: d=768, 12 heads, G=4
]


@pytest.mark.parametrize("B,N,n_fft,C,dg,with_mem", SWEEP)
def test_sweep_fp32(B, N, n_fft, C, dg, with_mem, fb, oracle, dev):
    V, gate, mem = _rand_case(B, N, n_fft, C, dg, with_mem, seed=B * 1000 + N + C)
    want = oracle.mix_flat(V, gate, n_fft, dg, mem)
    got = fb.spectral_mix(V.to(dev), gate.to(dev), None if mem is None else mem.to(dev), n_fft=n_fft, group_width=dg)
    _check(got, want.numpy())


@pytest.mark.parametrize("B,N,n_fft,C,dg,with_mem", [
    (2, 128, 128, 64, 4, True), (2, 1024, 1024, 64, 16, False), (1, 4096, 4096, 96, 16, True),
    (2, 3000, 4096, 32, 16, False), (2, 256, 256, 36, 6, True), (2, 256, 256, 15, 3, False),
    (1, 16384, 16384, 16, 16, False),
])
def test_sweep_bf16(B, N, n_fft, C, dg, with_mem, fb, oracle, dev):
    V, gate, mem = _rand_case(B, N, n_fft, C, dg, with_mem, seed=7 + N + C)
    Vb = V.to(torch.bfloat16)
    want = oracle.mix_flat(Vb.float(), gate, n_fft, dg, mem)     # oracle on the bf16-rounded input
    got = fb.spectral_mix(Vb.to(dev), gate.to(dev), None if mem is None else mem.to(dev), n_fft=n_fft, group_width=dg)
    assert got.dtype == torch.bfloat16
    _check(got, want.numpy(), rl2=REL_L2_BF16, mabs=2e-2)


def test_baseline_config2_full_size(fb, oracle, dev):
    """BASELINE config 2: batch=32 seq=1024 d=768 fp32, compared element-wise with the oracle."""
    V, gate, _ = _rand_case(32, 1024, 1024, 768, 16, False, seed=2)
    want = oracle.mix_flat(V, gate, 1024, 16)
    got = fb.spectral_mix(V.to(dev), gate.to(dev), n_fft=1024, group_width=16)
    _check(got, want.numpy())


def test_metric_shape_seq4096_d768(fb, oracle, dev):
    """The metric's shape (seq=4096, d=768, 48 gate groups), B=4, with memory; oracle head loop as the reference runs it."""
    V, gate, mem = _rand_case(4, 4096, 4096, 768, 16, True, seed=3)
    want = oracle.mix_head_loop(V, gate, 4096, 12, mem)
    got = fb.spectral_mix(V.to(dev), gate.to(dev), mem.to(dev), n_fft=4096, group_width=16)
    _check(got, want.numpy())


def test_long_context_16384_d768(fb, oracle, dev):
    """BASELINE config 4: seq=16384 d=768."""
    V, gate, _ = _rand_case(1, 16384, 16384, 768, 16, False, seed=4)
    want = oracle.mix_flat(V, gate, 16384, 16)
    got = fb.spectral_mix(V.to(dev), gate.to(dev), n_fft=16384, group_width=16)
    _check(got, want.numpy())


# --------------------------------------------------------------------------- size-independent properties at full size
def test_properties_at_full_size(fb, dev):
    """B=16, seq=4096, d=768 on the device only: identity gate, linearity, shift, memory-only."""
    B, N, C, dg = 16, 4096, 768, 16
    gen = torch.Generator(device=dev).manual_seed(5)
    V1 = torch.randn(B, N, C, device=dev, generator=gen)
    V2 = torch.randn(B, N, C, device=dev, generator=gen)
    F_half = N // 2 + 1
    ones = torch.ones(B, C // dg, F_half, dtype=torch.cfloat, device=dev)
    y = fb.spectral_mix(V1, ones, n_fft=N, group_width=dg)
    assert (y - V1).abs().max().item() < 1e-4 * V1.abs().max().item()          # identity gate -> identity
    gate = torch.randn(B, C // dg, F_half, dtype=torch.cfloat, device=dev, generator=gen)
    ya = fb.spectral_mix(V1, gate, n_fft=N, group_width=dg)
    yb = fb.spectral_mix(V2, gate, n_fft=N, group_width=dg)
    yab = fb.spectral_mix(2.0 * V1 - 3.0 * V2, gate, n_fft=N, group_width=dg)
    lin = 2.0 * ya - 3.0 * yb
    assert (yab - lin).norm().item() <= 2e-6 * lin.norm().item() + 1e-6        # linearity in V
    s = 37
    k = torch.arange(F_half, device=dev)
    shift = torch.exp(-2j * torch.pi * k * s / N).to(torch.cfloat).expand(B, C // dg, -1).contiguous()
    ys = fb.spectral_mix(V1, shift, n_fft=N, group_width=dg)
    assert (ys - torch.roll(V1, s, dims=1)).abs().max().item() < 2e-4 * V1.abs().max().item()  # circular shift
    mem = torch.randn(F_half, C, dtype=torch.cfloat, device=dev, generator=gen)
    ym = fb.spectral_mix(V1, torch.zeros_like(gate), mem, n_fft=N, group_width=dg)
    assert (ym - ym[0:1]).abs().max().item() == 0.0                            # zero gate: every row is irfft(memory)
    # checksum of checksums: sum over time of the output equals DC bin algebra: sum_n y = Re(G0) * sum_n v + Re(M0)
    g0 = gate[:, :, 0].real.repeat_interleave(dg, dim=1)
    assert torch.allclose(ya.sum(1), g0 * V1.sum(1), rtol=1e-3, atol=5e-2)


def test_kat_impulse_response(fb, dev):
    """An impulse at n=0 returns irfft(gate) per group (SURVEY section 7 KAT list)."""
    n_fft, C, dg = 1024, 32, 8
    V = torch.zeros(1, n_fft, C, device=dev)
    V[:, 0, :] = 1.0
    gate = torch.randn(1, C // dg, n_fft // 2 + 1, dtype=torch.cfloat, generator=torch.Generator().manual_seed(9))
    want = torch.fft.irfft(gate[0], n=n_fft, dim=-1).t().repeat_interleave(dg, dim=1)   # (n_fft, C) on CPU
    got = fb.spectral_mix(V, gate.to(dev), n_fft=n_fft, group_width=dg)
    _check(got[0], want.numpy())


def test_dc_nyquist_imag_ignored(fb, dev):
    V, gate, mem = _rand_case(2, 256, 256, 16, 4, True, seed=11)
    y0 = fb.spectral_mix(V.to(dev), gate.to(dev), mem.to(dev), n_fft=256, group_width=4)
    gate2, mem2 = gate.clone(), mem.clone()
    gate2[:, :, 0] = torch.complex(gate[:, :, 0].real, torch.randn(2, 4))
    mem2[-1] = torch.complex(mem[-1].real, torch.randn(16))
    mem2[0] = torch.complex(mem[0].real, torch.randn(16))
    # gate imag at DC multiplies a real bin -> contributes only to the ignored imaginary part
    y1 = fb.spectral_mix(V.to(dev), gate2.to(dev), mem2.to(dev), n_fft=256, group_width=4)
    assert torch.equal(y0, y1)


def test_tma_and_direct_load_paths_agree(fb, oracle, dev):
    """The TMA-staged tile load and the direct 128-bit global load feed the same arithmetic: bitwise equal."""
    from fft_b200 import _lib
    lib = _lib.load()
    for (B, N, n_fft, C, dg) in [(3, 4096, 4096, 64, 16), (40, 3500, 4096, 40, 8), (2, 900, 1024, 40, 8), (2, 2048, 2048, 32, 16)]:
        V, gate, mem = _rand_case(B, N, n_fft, C, dg, True, seed=21 + N)
        args = (V.to(dev), gate.to(dev), mem.to(dev))
        try:
            lib.spectre_mix_set_tma(0)
            y0 = fb.spectral_mix(*args, n_fft=n_fft, group_width=dg)
            lib.spectre_mix_set_tma(1)
            lib.spectre_mix_set_tmem(0)
            y1 = fb.spectral_mix(*args, n_fft=n_fft, group_width=dg)
            lib.spectre_mix_set_prefetch(1)
            y2 = fb.spectral_mix(*args, n_fft=n_fft, group_width=dg)
            lib.spectre_mix_set_tmem(1)          # tile I/O staged through tensor memory where a variant exists (4096 fp32)
            y3 = fb.spectral_mix(*args, n_fft=n_fft, group_width=dg)
            lib.spectre_mix_set_prefetch(0)
            lib.spectre_mix_set_skew_ns(0)       # no warp stagger, plain CTA barrier around the last pass's read
            lib.spectre_mix_set_sched(0)
            y4 = fb.spectral_mix(*args, n_fft=n_fft, group_width=dg)
        finally:
            lib.spectre_mix_set_tma(1)
            lib.spectre_mix_set_tmem(1)
            lib.spectre_mix_set_prefetch(0)
            lib.spectre_mix_set_skew_ns(-350)
            lib.spectre_mix_set_sched(3)
        y5 = fb.spectral_mix(*args, n_fft=n_fft, group_width=dg)   # defaults again
        assert torch.equal(y0, y1) and torch.equal(y1, y2) and torch.equal(y1, y3) and torch.equal(y1, y4) and torch.equal(y1, y5)
        _check(y1, oracle.mix_flat(V, gate, n_fft, dg, mem).numpy())


# --------------------------------------------------------------------------- layouts
def test_strided_views(fb, oracle, dev):
    """V as a channel slice of a wider tensor, memory as a chunk view (row stride > C), as spectre.py:703-707 makes them."""
    gen = torch.Generator().manual_seed(12)
    wide = torch.randn(2, 512, 96, generator=gen)
    memw = torch.randn(257, 96, dtype=torch.cfloat, generator=gen)
    gate = torch.randn(2, 4, 257, dtype=torch.cfloat, generator=gen)
    V, mem = wide[:, :, 32:64], memw[:, 32:64]
    want = oracle.mix_flat(V, gate, 512, 8, mem)
    dV, dM = wide.to(dev)[:, :, 32:64], memw.to(dev)[:, 32:64]
    assert not dV.is_contiguous()
    _check(fb.spectral_mix(dV, gate.to(dev), dM, n_fft=512, group_width=8), want.numpy())
    # odd channel offset forces the narrower element modes
    V2 = wide[:, :, 1:33]
    want2 = oracle.mix_flat(V2, gate, 512, 8)
    _check(fb.spectral_mix(wide.to(dev)[:, :, 1:33], gate.to(dev), n_fft=512, group_width=8), want2.numpy())


def test_empty_and_errors(fb, dev):
    gate = torch.ones(0, 2, 65, dtype=torch.cfloat, device=dev)
    y = fb.spectral_mix(torch.zeros(0, 128, 8, device=dev), gate, n_fft=128, group_width=4)
    assert y.shape == (0, 128, 8)
    with pytest.raises(RuntimeError, match="power of two"):
        fb.spectral_mix(torch.zeros(1, 100, 8, device=dev), torch.ones(1, 2, 51, dtype=torch.cfloat, device=dev),
                        n_fft=100, group_width=4)
    with pytest.raises(ValueError):
        fb.spectral_mix(torch.zeros(1, 128, 8, device=dev), torch.ones(1, 2, 64, dtype=torch.cfloat, device=dev),
                        n_fft=128, group_width=4)


# --------------------------------------------------------------------------- the other entry points
def test_rfft_seq(fb, dev):
    gen = torch.Generator().manual_seed(13)
    for (B, N, n_fft, C) in [(2, 128, 128, 16), (1, 200, 256, 32), (2, 1024, 1024, 12), (1, 3000, 4096, 8)]:
        V = torch.randn(B, N, C, generator=gen)
        want = torch.fft.rfft(V, n=n_fft, dim=1)
        got = fb.rfft_seq(V.to(dev), n_fft).cpu()
        assert got.shape == want.shape
        assert (got - want).norm() / want.norm() < 1e-6
    g = load_golden("decode_prefill_n256_d32")          # PrefixFFTCache.prefill of the reference (spectre.py:776-777)
    got = fb.rfft_seq(torch.from_numpy(g["V"]).to(dev), 256).cpu().numpy()
    assert rel_l2(got.view(np.float32), g["prefix_fft"].view(np.float32)) < 1e-6


def test_host_entry_point(fb, oracle):
    V, gate, mem = _rand_case(5, 1024, 1024, 64, 16, True, seed=14)
    want = oracle.mix_flat(V, gate, 1024, 16, mem)
    got = fb.spectral_mix_host(V, gate, mem, n_fft=1024, group_width=16)
    _check(got, want.numpy())
    got2 = fb.spectral_mix_host(V.pin_memory(), gate.pin_memory(), None, n_fft=1024, group_width=16)
    _check(got2, oracle.mix_flat(V, gate, 1024, 16).numpy())


def test_host_entry_ramped_chunk_schedule(fb, dev):
    """spectre_mix_fwd_host walks the batch in chunks that ramp up from ~12 MB to ~100 MB and back down (short pipeline fill and
    drain, few hand-overs in between): enough rows for the full ramp, ragged N, checked bit for bit against the device-buffer op
    on the same rows (which the other tests tie to the oracle)."""
    torch.manual_seed(77)
    B, N, n_fft, C, dg = 40, 3900, 4096, 768, 16          # 11.98 MB per row: chunks of 1, 2, 4, 8, 8, 8, 2+..., 4, 2, 1 rows
    V = torch.randn(B, N, C).pin_memory()
    gate = torch.randn(B, C // dg, n_fft // 2 + 1, dtype=torch.cfloat).pin_memory()
    got = fb.spectral_mix_host(V, gate, None, n_fft=n_fft, group_width=dg)
    want = fb.spectral_mix(V.to(dev), gate.to(dev), n_fft=n_fft, group_width=dg).cpu()
    assert got.shape == want.shape == (B, N, C)
    assert torch.equal(got, want)
    # the staging buffers are kept between calls; releasing them returns the memory and the next call allocates again
    from fft_b200 import _lib
    lib = _lib.load()
    torch.cuda.synchronize()
    free0 = torch.cuda.mem_get_info()[0]
    assert lib.spectre_mix_host_release() == 0
    assert torch.cuda.mem_get_info()[0] - free0 > 300 << 20
    assert lib.spectre_mix_host_release() == 0                       # idempotent
    got2 = fb.spectral_mix_host(V[:3], gate[:3], None, n_fft=n_fft, group_width=dg)
    assert torch.equal(got2, want[:3])


@pytest.mark.parametrize("n_fft,B,C,dg", [(8192, 5, 64, 16), (16384, 4, 64, 16), (8192, 4, 768, 16), (16384, 4, 768, 16)])
def test_host_entry_point_long_context(n_fft, B, C, dg, fb, oracle):
    """spectre_mix_fwd_host at n_fft = 8192 / 16384 with several chunks in flight (one batch row per chunk at C = 768: the
    chunks run on 4 streams, each with its OWN workspace -- the round-1 per-device scratch raced here)."""
    V, gate, mem = _rand_case(B, n_fft - 100, n_fft, C, dg, C == 64, seed=140 + B + C)
    want = oracle.mix_flat(V, gate, n_fft, dg, mem).numpy()
    import os
    old = os.environ.get("SPECTRE_MIX_HOST_CHUNK_MB")
    if C == 64:
        os.environ["SPECTRE_MIX_HOST_CHUNK_MB"] = "4"       # force one row per chunk at the narrow shape too
    try:
        got = fb.spectral_mix_host(V, gate, mem, n_fft=n_fft, group_width=dg)
    finally:
        if old is None:
            os.environ.pop("SPECTRE_MIX_HOST_CHUNK_MB", None)
        else:
            os.environ["SPECTRE_MIX_HOST_CHUNK_MB"] = old
    _check(got, want)
    with pytest.raises(ValueError):
        fb.spectral_mix_host(V, gate, mem, n_fft=n_fft, group_width=dg, out=torch.empty(B, 8, C))


def test_long_context_concurrent_streams_and_workspace(fb, oracle, dev):
    """Two side streams run the n_fft = 16384 path at the same time (different inputs): results are bit-equal to the serial
    ones, through the torch op (workspace from the caching allocator), through the plain C entry without a workspace
    (stream-ordered pool allocation per call) and through spectre_mix_fwd_ws with caller-owned scratch."""
    import ctypes
    from fft_b200 import _lib
    lib = _lib.load()
    n_fft, C, dg, B = 16384, 256, 16, 6
    ins = []
    for i in range(2):
        V, gate, _ = _rand_case(B, n_fft, n_fft, C, dg, False, seed=170 + i)
        ins.append((V.to(dev), gate.to(dev)))
    serial = [fb.spectral_mix(V, g, n_fft=n_fft, group_width=dg).clone() for V, g in ins]
    _check(serial[0][:1], oracle.mix_flat(ins[0][0][:1].cpu(), ins[0][1][:1].cpu(), n_fft, dg).numpy())
    need = lib.spectre_mix_workspace_bytes(0, B, n_fft, n_fft, C, dg)
    assert need == B * n_fft * C * 4 and fb.plan_info(B, n_fft, n_fft, C, dg)["workspace_bytes"] == need
    assert lib.spectre_mix_workspace_bytes(0, B, 4096, 4096, C, dg) == 0

    def call(kind, V, g, out, ws, stream):
        args = (V.data_ptr(), 0, V.stride(0), V.stride(1), g.data_ptr(), None, 0, out.data_ptr(), 0, out.stride(0), out.stride(1),
                B, n_fft, n_fft, C, dg)
        if kind == "plain":
            rc = lib.spectre_mix_fwd(*args, ctypes.c_void_p(stream.cuda_stream))
        else:
            rc = lib.spectre_mix_fwd_ws(*args, ws.data_ptr(), ws.numel(), ctypes.c_void_p(stream.cuda_stream))
        _lib.check(rc, kind)

    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    for kind in ("op", "plain", "ws"):
        for rep in range(3):
            outs = [torch.empty_like(serial[0]) for _ in range(2)]
            wss = [torch.empty(need, dtype=torch.uint8, device=dev) for _ in range(2)]
            torch.cuda.synchronize()
            for i, st in enumerate(streams):
                with torch.cuda.stream(st):
                    for _ in range(2):      # back to back on each stream as well
                        if kind == "op":
                            outs[i] = fb.spectral_mix(*ins[i], n_fft=n_fft, group_width=dg)
                        else:
                            call(kind, ins[i][0], ins[i][1], outs[i], wss[i], st)
            torch.cuda.synchronize()
            for i in range(2):
                assert torch.equal(outs[i], serial[i]), (kind, rep, i)
    # a workspace that is too small is refused, not overrun
    small = torch.empty(need // 2, dtype=torch.uint8, device=dev)
    rc = lib.spectre_mix_fwd_ws(ins[0][0].data_ptr(), 0, ins[0][0].stride(0), ins[0][0].stride(1), ins[0][1].data_ptr(), None, 0,
                                outs[0].data_ptr(), 0, outs[0].stride(0), outs[0].stride(1), B, n_fft, n_fft, C, dg,
                                small.data_ptr(), small.numel(), None)
    assert rc == 1 and b"workspace too small" in lib.spectre_mix_last_error()


def test_cuda_graph_capture_long_context(fb, dev):
    """n_fft = 16384 (three launches around a workspace) captured into a CUDA graph and replayed on new data: through the
    torch op and through the plain C entry (stream-ordered allocation inside the capture)."""
    import ctypes
    from fft_b200 import _lib
    lib = _lib.load()
    n_fft, C, dg, B = 16384, 64, 16, 3
    V, gate, _ = _rand_case(B, n_fft, n_fft, C, dg, False, seed=178)
    Vd, gd = V.to(dev), gate.to(dev)
    ref1 = fb.spectral_mix(Vd, gd, n_fft=n_fft, group_width=dg).clone()          # warm-up (twiddle tables)
    ref2 = fb.spectral_mix(2.0 * Vd, gd, n_fft=n_fft, group_width=dg).clone()
    static_in = Vd.clone()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        static_out = fb.spectral_mix(static_in, gd, n_fft=n_fft, group_width=dg)
    graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(static_out, ref1)
    static_in.copy_(2.0 * Vd)
    graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(static_out, ref2)
    out2 = torch.empty_like(ref1)
    g2 = torch.cuda.CUDAGraph()
    static_in.copy_(Vd)
    # one warm-up call of the plain entry outside the capture (creates the device's private pool)
    _lib.check(lib.spectre_mix_fwd(static_in.data_ptr(), 0, static_in.stride(0), static_in.stride(1), gd.data_ptr(), None, 0,
                                   out2.data_ptr(), 0, out2.stride(0), out2.stride(1), B, n_fft, n_fft, C, dg,
                                   ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)), "warm-up")
    torch.cuda.synchronize()
    with torch.cuda.graph(g2):
        rc = lib.spectre_mix_fwd(static_in.data_ptr(), 0, static_in.stride(0), static_in.stride(1), gd.data_ptr(), None, 0,
                                 out2.data_ptr(), 0, out2.stride(0), out2.stride(1), B, n_fft, n_fft, C, dg,
                                 ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
        _lib.check(rc, "spectre_mix_fwd under capture")
    g2.replay()
    torch.cuda.synchronize()
    assert torch.equal(out2, ref1)
    static_in.copy_(2.0 * Vd)
    g2.replay()
    torch.cuda.synchronize()
    assert torch.equal(out2, ref2)


def test_plan_info(fb):
    info = fb.plan_info(8, 4096, 4096, 768, 16)
    assert info["n_fft"] == 4096 and np.prod(info["radix"]) == 4096
    assert info["algorithmic_bytes"] == 8 * 4096 * 768 * 8 + 8 * 48 * 2049 * 8   # 6336.1 B/token
    assert info["launches"] == 1 and info["grid"] >= 1


def test_plan_is_batch_aware_for_wide_row_variants(fb):
    """seq 1024 / 2048: the wide-row TMEM-staged variants (32- / 16-channel tiles, one CTA per SM) from about seven tiles per SM
    on, the narrower two-CTAs-per-SM variants for shorter launches."""
    assert fb.plan_info(256, 1024, 1024, 768, 16)["tile_channels"] == 32 and fb.plan_info(256, 1024, 1024, 768, 16)["radix"] == [16, 16, 4]
    assert fb.plan_info(32, 1024, 1024, 768, 16)["tile_channels"] == 16 and fb.plan_info(32, 1024, 1024, 768, 16)["radix"] == [4, 16, 16]
    assert fb.plan_info(128, 2048, 2048, 768, 16)["tile_channels"] == 16 and fb.plan_info(128, 2048, 2048, 768, 16)["radix"] == [16, 16, 8]
    assert fb.plan_info(8, 2048, 2048, 768, 16)["tile_channels"] == 8


def test_wide_row_variants_against_oracle(fb, oracle, dev):
    """The wide-row variants at batch sizes that select them, with memory, ragged N, ragged channel tiles (C not a multiple of the
    tile width), group widths 8 / 16 / 32 (4, 2 and 1 gate tables per 32-channel tile) and bf16."""
    for (B, N, n_fft, C, dg, with_mem) in [(44, 1024, 1024, 768, 16, False), (12, 1000, 1024, 3112, 8, True),
                                           (32, 1024, 1024, 1056, 32, False), (22, 2048, 2048, 784, 16, True),
                                           (11, 2000, 2048, 1608, 8, False)]:
        V, gate, mem = _rand_case(B, N, n_fft, C, dg, with_mem, seed=400 + C)
        info = fb.plan_info(B, N, n_fft, C, dg)
        assert info["radix"][0] == 16 and info["radix"][-1] in (4, 8), info          # the wide-row plan was chosen
        got = fb.spectral_mix(V.to(dev), gate.to(dev), None if mem is None else mem.to(dev), n_fft=n_fft, group_width=dg)
        _check(got[:3], oracle.mix_flat(V[:3], gate[:3], n_fft, dg, mem).numpy())
        _check(got[-2:], oracle.mix_flat(V[-2:], gate[-2:], n_fft, dg, mem).numpy())
    V, gate, _ = _rand_case(50, 1024, 1024, 768, 16, False, seed=410)
    got = fb.spectral_mix(V.to(dev).to(torch.bfloat16), gate.to(dev), n_fft=1024, group_width=16)
    _check(got[:2], oracle.mix_flat(V[:2].to(torch.bfloat16).float(), gate[:2], 1024, 16).numpy(), rl2=REL_L2_BF16, mabs=2e-2)


def test_autograd_matches_oracle(fb, oracle, dev):
    """Backward of the op against autograd through the oracle's torch.fft path."""
    for (B, N, n_fft, C, dg) in [(2, 128, 128, 16, 4), (2, 100, 128, 16, 8), (1, 1024, 1024, 16, 16)]:
        V, gate, mem = _rand_case(B, N, n_fft, C, dg, True, seed=15 + N)
        w = torch.randn(B, min(N, n_fft), C, generator=torch.Generator().manual_seed(1))
        Vc, gc, mc = V.clone().requires_grad_(), gate.clone().requires_grad_(), mem.clone().requires_grad_()
        (oracle.mix_flat(Vc, gc, n_fft, dg, mc) * w).sum().backward()
        Vg, gg, mg = (V.to(dev).requires_grad_(), gate.to(dev).requires_grad_(), mem.to(dev).requires_grad_())
        (fb.spectral_mix(Vg, gg, mg, n_fft=n_fft, group_width=dg) * w.to(dev)).sum().backward()
        assert rel_l2(Vg.grad.cpu().numpy(), Vc.grad.numpy()) < 1e-5
        assert rel_l2(torch.view_as_real(gg.grad).cpu().numpy(), torch.view_as_real(gc.grad).numpy()) < 1e-5
        assert rel_l2(torch.view_as_real(mg.grad).cpu().numpy(), torch.view_as_real(mc.grad).numpy()) < 1e-5


def test_autograd_at_metric_shape_fused_gate_gradient(fb, oracle, dev):
    """Backward at the metric shape (2, 4096, 768), 48 gate groups: dV through the mix kernel, dgate through the fused
    gate-gradient kernel (spectre_mix_dgate: both transforms and the group reduction in one launch), dmemory by linearity;
    against autograd through the oracle's torch.fft path.  Also ragged N, group widths 8 / 24 and bf16 inputs."""
    import ctypes
    from fft_b200 import _lib, ops
    lib = _lib.load()
    for (B, N, n_fft, C, dg, dt) in [(2, 4096, 4096, 768, 16, torch.float32), (2, 3000, 4096, 48, 8, torch.float32),
                                     (1, 4096, 4096, 48, 24, torch.float32), (2, 4096, 4096, 64, 16, torch.bfloat16)]:
        V, gate, mem = _rand_case(B, N, n_fft, C, dg, True, seed=300 + C + N)
        if dt == torch.bfloat16:
            V = V.to(dt).float()
        w = torch.randn(B, min(N, n_fft), C, generator=torch.Generator().manual_seed(2))
        Vc, gc, mc = V.clone().requires_grad_(), gate.clone().requires_grad_(), mem.clone().requires_grad_()
        (oracle.mix_flat(Vc, gc, n_fft, dg, mc) * w).sum().backward()
        # the fused kernel is the one that runs here
        dg_direct = ops._dgate_fused(V.to(dev).to(dt), w.to(dev).to(dt), n_fft, dg)
        assert dg_direct is not None, "fused gate-gradient kernel not used at n_fft = 4096"
        tol = 1e-5 if dt == torch.float32 else 1e-2
        assert rel_l2(torch.view_as_real(dg_direct).cpu().numpy(), torch.view_as_real(gc.grad).numpy()) < tol
        if dt == torch.float32:
            Vg, gg, mg = (V.to(dev).requires_grad_(), gate.to(dev).requires_grad_(), mem.to(dev).requires_grad_())
            (fb.spectral_mix(Vg, gg, mg, n_fft=n_fft, group_width=dg) * w.to(dev)).sum().backward()
            assert rel_l2(Vg.grad.cpu().numpy(), Vc.grad.numpy()) < 1e-5
            assert rel_l2(torch.view_as_real(gg.grad).cpu().numpy(), torch.view_as_real(gc.grad).numpy()) < 1e-5
            assert rel_l2(torch.view_as_real(mg.grad).cpu().numpy(), torch.view_as_real(mc.grad).numpy()) < 1e-5
    # imaginary parts at DC / Nyquist are exactly zero, and the call is repeatable bit for bit at group width 16 (two tiles per group)
    a = ops._dgate_fused(V.to(dev).to(dt), w.to(dev).to(dt), n_fft, dg)
    b = ops._dgate_fused(V.to(dev).to(dt), w.to(dev).to(dt), n_fft, dg)
    assert torch.equal(a, b) and float(a.imag[:, :, 0].abs().max()) == 0.0 and float(a.imag[:, :, -1].abs().max()) == 0.0
    # sizes without a fused variant report UNSUPPORTED (the binding then takes the two-spectrum route, tested above at 128 / 1024)
    x = torch.randn(1, 1024, 16, device=dev)
    out = torch.empty(1, 1, 513, dtype=torch.complex64, device=dev)
    rc = lib.spectre_mix_dgate(x.data_ptr(), x.data_ptr(), 0, x.stride(0), x.stride(1), x.stride(0), x.stride(1), out.data_ptr(),
                               1, 1024, 1024, 16, 16, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 2 and b"no fused gate-gradient kernel" in lib.spectre_mix_last_error()


# --------------------------------------------------------------------------- drop-in module shells
@pytest.mark.parametrize("name", ["block_d64_h4_n128_mem", "block_d64_h4_n128"])
def test_block_shell_matches_reference_block(name, fb, dev):
    """Stock reference SpectreBlock output (golden) vs our shell loaded from the reference's state_dict."""
    g = load_golden(name)
    blk = fb.SpectreBlock(64, 4, 128, pooling_type="mean", wavelet_on_rate=0.0, memory_size=int(g["memory_size"]))
    sd = {k[4:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("sd::")}
    blk.load_state_dict(sd, strict=True)
    blk = blk.to(dev).eval()
    with torch.no_grad():
        y = blk(torch.from_numpy(g["x"]).to(dev))
    _check(y, g["y"], rl2=2e-5, mabs=2e-4)   # includes the stock fp32 GEMMs/LayerNorms run on the GPU


def test_multihead_matches_per_head_calls(fb, dev):
    """One fused launch over all heads == the reference's per-head loop (spectre.py:712-718)."""
    torch.manual_seed(16)
    mh = fb.SpectreMultiHead(64, 4, 256, pooling_type="mean", wavelet_on_rate=0.0).to(dev).eval()
    x = torch.randn(2, 200, 64, device=dev)
    mem = torch.randn(129, 64, dtype=torch.cfloat, device=dev)
    with torch.no_grad():
        fused = mh(x, memory_fft=mem)
        per_head = [h(c, None, return_q_pool=True, memory_fft=m)[0]
                    for h, c, m in zip(mh.heads, torch.chunk(x, 4, -1), torch.chunk(mem, 4, -1))]
        loop = mh.out_proj(torch.cat(per_head, dim=-1))
    assert (fused - loop).norm() / loop.norm() < 1e-5


# --------------------------------------------------------------------------- gate generator tail
@pytest.mark.parametrize("name", golden_names("gate_"))
def test_gate_expand_matches_reference_goldens(name, fb, dev):
    """Cubic interpolation + modReLU + positional phase in one launch vs the reference's own functions (goldens)."""
    g = load_golden(name)
    F_half = int(g["n_fft"]) // 2 + 1
    a = torch.from_numpy(g["anchors"]).to(dev)
    G = a.shape[1]
    bias = torch.from_numpy(g["bias"]).to(dev).view(G, F_half)
    eps = torch.from_numpy(g["eps"]).to(dev).reshape(1).expand(G)
    for key, pos in (("gate", None), ("gate_pos1", g["pos1"]), ("gate_posB", g["posB"]), ("gate_pos1", g["pos1"][0])):
        p = None if pos is None else torch.from_numpy(pos).to(dev)
        got = torch.view_as_real(fb.gate_expand(a, bias, eps, p, F_half=F_half, G=G))
        _check(got, torch.view_as_real(torch.from_numpy(g[key])).numpy(), rl2=5e-6, mabs=2e-5)
    # several heads stacked: the row shuffle of spectre.py:41 stays inside each head
    a2 = torch.cat([a, a.flip(1)], dim=1)
    got = fb.gate_expand(a2, torch.cat([bias, bias]), torch.cat([eps, eps]), None, F_half=F_half, G=G)
    _check(torch.view_as_real(got[:, :G]), torch.view_as_real(torch.from_numpy(g["gate"])).numpy(), rl2=5e-6, mabs=2e-5)


@pytest.mark.parametrize("name", ["gate_n128_g4", "gate_n4096_g4"])
def test_mix_with_fused_gate_generator_matches_goldens(name, fb, oracle, dev):
    """spectre_mix_fwd_anchors: anchors -> [cubic interpolation, modReLU, phase inside the mix kernel] -> mix in ONE
    launch, against the reference-generated gate goldens composed with the oracle's mix; packed layout (fused), a layout that
    has no fused variant (group width 6: gate expanded into the workspace), memory, several heads, bf16."""
    g = load_golden(name)
    n_fft = int(g["n_fft"])
    F_half = n_fft // 2 + 1
    a = torch.from_numpy(g["anchors"])
    B, G, _ = a.shape
    bias = torch.from_numpy(g["bias"]).view(G, F_half)
    eps = torch.from_numpy(g["eps"]).reshape(1).expand(G)
    gen = torch.Generator().manual_seed(90)
    for dg, with_mem, N in [(16, False, n_fft), (16, True, n_fft - 28), (8, False, n_fft), (6, True, n_fft)]:
        C = G * dg
        V = torch.randn(B, N, C, generator=gen)
        mem = torch.randn(F_half, C, dtype=torch.cfloat, generator=gen) / 8 if with_mem else None
        for key, pos in (("gate", None), ("gate_pos1", g["pos1"]), ("gate_posB", g["posB"])):
            want = oracle.mix_flat(V, torch.from_numpy(g[key]), n_fft, dg, mem).numpy()
            p = None if pos is None else torch.from_numpy(pos).to(dev)
            got = fb.spectral_mix_anchors(V.to(dev), a.to(dev), bias.to(dev), eps.to(dev), p, None if mem is None else mem.to(dev),
                                          n_fft=n_fft, group_width=dg, G=G)
            _check(got, want)
    # two heads stacked (NG = 2 G, the second head's anchor rows flipped): the row shuffle of spectre.py:41 stays inside a head
    a2 = torch.cat([a, a.flip(1)], dim=1).to(dev)
    V2 = torch.randn(B, n_fft, 2 * G * 16, generator=gen).to(dev)
    b2, e2 = torch.cat([bias, bias]).to(dev), torch.cat([eps, eps]).to(dev)
    fused = fb.spectral_mix_anchors(V2, a2, b2, e2, n_fft=n_fft, group_width=16, G=G)
    unfused = fb.spectral_mix(V2, fb.gate_expand(a2, b2, e2, None, F_half=F_half, G=G), n_fft=n_fft, group_width=16)
    _check(fused, unfused.cpu().numpy(), rl2=1e-6, mabs=1e-5)
    fb16 = fb.spectral_mix_anchors(V2.to(torch.bfloat16), a2, b2, e2, n_fft=n_fft, group_width=16, G=G)
    _check(fb16, unfused.cpu().numpy(), rl2=REL_L2_BF16, mabs=2e-2)


def test_mix_with_fused_gate_generator_other_sizes_and_autograd(fb, dev):
    """Fused gate generator at every kernel family (small, 1024, 2048 TMA variants; 8192 / 16384 two-pass sub-transform
    kernel) against gate_expand + spectral_mix, and its backward (V, anchors, bias, memory) against the unfused composition."""
    gen = torch.Generator().manual_seed(91)
    for (B, N, n_fft, H, G, dg) in [(2, 64, 64, 2, 4, 4), (3, 1000, 1024, 3, 4, 16), (2, 2048, 2048, 2, 2, 8),
                                    (2, 8192, 8192, 1, 4, 16), (1, 16000, 16384, 2, 4, 8)]:
        F_half, NG, C = n_fft // 2 + 1, H * G, H * G * dg
        Bk = max(4, int(F_half ** 0.5))
        a = torch.randn(B, NG, Bk, dtype=torch.cfloat, generator=gen).to(dev)
        bias = (0.3 * torch.randn(NG, F_half, generator=gen) - 0.1).to(dev)
        eps = torch.full((NG,), 1e-4).to(dev)
        V = torch.randn(B, N, C, generator=gen).to(dev)
        fused = fb.spectral_mix_anchors(V, a, bias, eps, n_fft=n_fft, group_width=dg, G=G)
        unfused = fb.spectral_mix(V, fb.gate_expand(a, bias, eps, None, F_half=F_half, G=G), n_fft=n_fft, group_width=dg)
        _check(fused, unfused.cpu().numpy(), rl2=2e-6, mabs=2e-5)
    # backward
    B, N, n_fft, G, dg = 2, 256, 256, 4, 8
    F_half, C = 129, 32
    a = torch.randn(B, G, 11, dtype=torch.cfloat, generator=gen).to(dev)
    bias = (0.3 * torch.randn(G, F_half, generator=gen)).to(dev)
    eps = torch.full((G,), 1e-4).to(dev)
    V = torch.randn(B, N, C, generator=gen).to(dev)
    mem = (torch.randn(F_half, C, dtype=torch.cfloat, generator=gen) / 8).to(dev)
    w = torch.randn(B, N, C, generator=gen).to(dev)
    grads = []
    for fused in (True, False):
        Vg, ag, bg, mg = (t.clone().requires_grad_() for t in (V, a, bias, mem))
        if fused:
            y = fb.spectral_mix_anchors(Vg, ag, bg, eps, None, mg, n_fft=n_fft, group_width=dg, G=G)
        else:
            y = fb.spectral_mix(Vg, fb.gate_expand(ag, bg, eps, None, F_half=F_half, G=G), mg, n_fft=n_fft, group_width=dg)
        (y * w).sum().backward()
        grads.append([Vg.grad, torch.view_as_real(ag.grad), bg.grad, torch.view_as_real(mg.grad)])
    for x, y in zip(*grads):
        assert rel_l2(x.cpu().numpy(), y.cpu().numpy()) < 1e-5


def test_block_forward_launches_no_gate_kernel(fb, dev):
    """The module path runs the gate generator's tail inside the mix kernel: a SpectreBlock forward launches the mix kernel
    and no stand-alone gate-expansion kernel."""
    from torch.profiler import ProfilerActivity, profile
    torch.manual_seed(92)
    blk = fb.SpectreBlock(64, 4, 128, pooling_type="mean", wavelet_on_rate=0.0).to(dev).eval()
    x = torch.randn(2, 128, 64, device=dev)
    with torch.no_grad():
        blk(x)
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            blk(x)
            torch.cuda.synchronize()
    names = [e.key for e in prof.key_averages()]
    assert any("spectre_mix_kernel" in n for n in names), names
    assert not any("gate_expand" in n for n in names), names


def test_gate_expand_autograd(fb, oracle, dev):
    torch.manual_seed(40)
    a = torch.randn(2, 8, 11, dtype=torch.cfloat)
    bias = -0.1 + 0.3 * torch.randn(8, 65)
    eps = torch.full((8,), 1e-4)
    w = torch.randn(2, 8, 65, dtype=torch.cfloat)
    ac, bc = a.clone().requires_grad_(True), bias.clone().requires_grad_(True)
    ref = torch.cat([oracle.gate_tail(ac[:, i:i + 4], bc[i:i + 4], eps[i:i + 4], None, 65) for i in (0, 4)], dim=1)
    (ref * w).real.sum().backward()
    ag, bg = a.to(dev).requires_grad_(True), bias.to(dev).requires_grad_(True)
    got = fb.gate_expand(ag, bg, eps.to(dev), None, F_half=65, G=4)
    (got * w.to(dev)).real.sum().backward()
    _check(torch.view_as_real(ag.grad), torch.view_as_real(ac.grad).numpy(), rl2=1e-5, mabs=1e-4)
    _check(bg.grad, bc.grad.numpy(), rl2=1e-5, mabs=1e-4)


def test_batched_gate_generator_matches_stock_per_head_ops(fb, dev):
    """All-heads gate generator (batched GEMMs + one expansion launch) vs the stock per-head op sequence of
    spectre.py:511-536 run on the same device."""
    import fft_b200.modules as M
    torch.manual_seed(41)
    mh = fb.SpectreMultiHead(768, 12, 1024, pooling_type="mean", wavelet_on_rate=0.0).to(dev).eval()
    x = torch.randn(2, 1024, 768, device=dev)
    with torch.no_grad():
        for h in mh.heads:
            h.modrelu.bias.add_(0.3 * torch.randn_like(h.modrelu.bias))
        Q_all = torch.stack([h.W_q(c) for h, c in zip(mh.heads, torch.chunk(x, 12, -1))], dim=2)
        gate, q_pool = M.multihead_gate(mh, Q_all, None)
        stock = []
        for i, h in enumerate(mh.heads):
            qp = h.q_norm(Q_all[:, :, i].mean(1))
            anc = torch.view_as_complex(h.gate_mlp(qp).view(2, h.G, h.B, 2))
            up = M.interp_complex_1d(anc, size=h.F_half, mode="cubic")
            stock.append(h.modrelu(up.reshape(2, -1)).view_as(up))
        stock = torch.cat(stock, dim=1)
    assert gate.shape == stock.shape == (2, 48, 513)
    _check(torch.view_as_real(gate), torch.view_as_real(stock).cpu().numpy(), rl2=2e-5, mabs=1e-4)


# --------------------------------------------------------------------------- decode side
def test_decode_sequence_matches_reference(fb, dev):
    """Prompt of 200 tokens then 100 single-token steps (eviction after t >= 256) against the outputs of the stock
    SpectreHead.decode_step / PrefixFFTCache recorded from the reference (spectre.py:562-611, :786-814)."""
    g = load_golden("decode_seq_n256_d32")
    head = fb.SpectreHead(32, 256, pooling_type="mean")
    head.load_state_dict({k[4:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("sd::")}, strict=True)
    head = head.to(dev).eval()
    cache = fb.PrefixFFTCache(256, 32, device=dev)
    cache.prefill(torch.from_numpy(g["Qp"]).to(dev), torch.from_numpy(g["Vp"]).to(dev))
    qs, vs = torch.from_numpy(g["qs"]).to(dev), torch.from_numpy(g["vs"]).to(dev)
    outs = torch.stack([head.decode_step(qs[i], vs[i], cache) for i in range(qs.shape[0])]).cpu().numpy()
    assert cache.t == int(g["t"])
    assert rel_l2(outs, g["outs"]) < 2e-5 and max_abs_rel(outs, g["outs"]) < 1e-4
    got = torch.view_as_real(cache.prefix_fft).cpu().numpy()
    want = np.ascontiguousarray(g["prefix_fft"]).view(np.float32).reshape(got.shape)
    assert rel_l2(got, want) < 1e-5
    assert rel_l2(cache.sum_q.cpu().numpy(), g["sum_q"]) < 1e-5


def test_decode_split_equals_fused(fb, dev):
    """cache.decode_step + cache.readout (two kernels) == the fused step; several heads share one cache (d = 64, d_g = 8)."""
    torch.manual_seed(30)
    n, d, dg = 128, 64, 8
    c1, c2 = fb.PrefixFFTCache(n, d, device=dev), fb.PrefixFFTCache(n, d, device=dev)
    Q, V = torch.randn(100, d, device=dev), torch.randn(100, d, device=dev)
    c1.prefill(Q, V)
    c2.prefill(Q, V)
    for step in range(60):
        q, v = torch.randn(d, device=dev), torch.randn(d, device=dev)
        gate = torch.randn(d // dg, n // 2 + 1, dtype=torch.cfloat, device=dev)
        c1.decode_step(q, v)
        a = c1.readout(gate, c1.t % n)
        j = c2._advance(q)
        b = c2.fused_step(v, c2.V_buf[j], gate)
        c2._store_v(j, v)
        assert torch.allclose(a, b, rtol=1e-5, atol=1e-5)
    assert torch.equal(c1.prefix_fft, c2.prefix_fft)


def test_decode_gate_kernel_matches_stock_ops_and_is_deterministic(fb, dev):
    """spectre_decode_gate (interp + modReLU + decode phase in one launch) against the stock PyTorch op sequence of
    spectre.py:578-598, before and after the window is full (phase angle in the reference's float32 rounding order);
    the read-out is bit-reproducible run to run (two-stage fixed-order reduction, no atomics)."""
    from fft_b200.decode import decode_gate, decode_gate_torch
    torch.manual_seed(31)
    n, d = 256, 32
    head = fb.SpectreHead(d, n, pooling_type="mean").to(dev).eval()
    cache = fb.PrefixFFTCache(n, d, device=dev)
    cache.prefill(torch.randn(200, d, device=dev), torch.randn(200, d, device=dev))
    for t in (210, 255, 256, 300, 1000, 5000):
        cache.t = t
        with torch.no_grad():
            a, b = decode_gate(head, cache), decode_gate_torch(head, cache)
        assert a.shape == b.shape == (head.G, head.F_half)
        assert rel_l2(torch.view_as_real(a).cpu().numpy(), torch.view_as_real(b).cpu().numpy()) < 2e-5, t
    cache.t = 230
    gate = torch.randn(d // 8, n // 2 + 1, dtype=torch.cfloat, device=dev)
    outs = [cache.readout(gate, 17).clone() for _ in range(5)]
    assert all(torch.equal(outs[0], o) for o in outs[1:])


# --------------------------------------------------------------------------- BASELINE config 3: full model, bf16
def test_spectre_base_bf16_forward(fb, dev):
    """12 blocks, d=768, 12 heads, seq=4096 under bf16 autocast: runs through the bf16 kernel variant and stays close
    to the same model evaluated in fp32 (the stock reference cannot run in bf16 on CUDA at all)."""
    torch.manual_seed(40)
    model = fb.SpectreBase(vocab=1000, depth=12).to(dev).eval()
    tokens = torch.randint(0, 1000, (1, 4096), device=dev)
    with torch.no_grad():
        h32 = model(tokens, return_hidden=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            h16 = model(tokens, return_hidden=True)
    assert h32.shape == (1, 4096, 768) and torch.isfinite(h16).all()
    err = (h16.float() - h32).norm() / h32.norm()
    assert err < 5e-2, err


def test_patch_reference_style_module(fb, dev):
    """patch_reference() rebinds forward on modules matched by class name; emulate a reference-built module tree."""
    import types
    torch.manual_seed(41)
    blk = fb.SpectreBlock(64, 4, 128, pooling_type="mean", wavelet_on_rate=0.0).to(dev).eval()
    x = torch.randn(2, 128, 64, device=dev)
    with torch.no_grad():
        y0 = blk(x)
    blk.mix.forward = types.MethodType(lambda self, *a, **k: (_ for _ in ()).throw(RuntimeError("stock forward")), blk.mix)
    assert fb.patch_reference(blk) >= 1
    with torch.no_grad():
        y1 = blk(x)
    assert torch.equal(y0, y1)


def _real_reference():
    """The UNMODIFIED reference module from the travelling copy baseline/_ref/spectre.py (made by build(); /root/reference does
    not exist on the GPU box)."""
    import importlib.util
    import os
    import warnings
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref", "spectre.py")
    if not os.path.exists(path):
        pytest.skip("baseline/_ref/spectre.py missing: run __graft_entry__.build() where /root/reference exists")
    spec = importlib.util.spec_from_file_location("_ref_spectre_t", path)
    ref = importlib.util.module_from_spec(spec)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        spec.loader.exec_module(ref)
    return ref


@pytest.mark.parametrize("name", ["block_d64_h4_n128_mem", "block_d64_h4_n128"])
def test_patch_real_reference_block(name, fb, dev):
    """A block built from the REAL reference classes (spectre.SpectreBlock, spectre.py:892-982), loaded with the golden's
    state_dict, moved to the GPU and switched with patch_reference(): reproduces the stock CPU output recorded in the golden."""
    ref = _real_reference()
    g = load_golden(name)
    blk = ref.SpectreBlock(64, 4, 128, pooling_type="mean", wavelet_on_rate=0.0, memory_size=int(g["memory_size"]))
    blk.load_state_dict({k[4:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("sd::")}, strict=True)
    x = torch.from_numpy(g["x"])
    with torch.no_grad():
        y_stock_cpu = blk(x)                                   # the reference itself, here, on the CPU
    assert rel_l2(y_stock_cpu.numpy(), g["y"]) < 1e-6
    blk = blk.to(dev).eval()
    assert fb.patch_reference(blk) >= 1 + 4                      # the SpectreMultiHead and its 4 heads
    with torch.no_grad():
        y = blk(x.to(dev))
    _check(y, g["y"], rl2=2e-5, mabs=2e-4)
    # a non-contiguous activation works like in the reference (torch.chunk accepts any strides, spectre.py:703)
    xt = torch.from_numpy(g["x"]).to(dev).transpose(0, 1).contiguous().transpose(0, 1)
    assert not xt.is_contiguous()
    with torch.no_grad():
        _check(blk.mix(blk.ln1(xt)), blk.mix(blk.ln1(xt.contiguous())).cpu().numpy(), rl2=1e-6, mabs=1e-5)


def test_spectre_base_fp32_matches_real_reference_stack(fb, dev):
    """BASELINE configs[2] tied to the reference: 12 x the REAL spectre.SpectreBlock(768, 12 heads, n_fft=4096) run on the
    host CPU (the stock code path: torch.fft / MKL) against the same weights in our SpectreBase on the GPU, fp32, seq 4096,
    batch 1, hidden states after the 12th block at rel-L2 <= 1e-4."""
    ref = _real_reference()
    torch.manual_seed(42)
    kw = dict(mlp_ratio=4, d_gate=256, pooling_type="mean", num_groups=4, wavelet_on_rate=0.0, memory_size=0)
    ref_blocks = [ref.SpectreBlock(768, 12, 4096, **kw).eval() for _ in range(12)]
    x = torch.randn(1, 4096, 768)
    with torch.no_grad():
        h = x
        for b in ref_blocks:
            h = b(h)
    model = fb.SpectreBase(vocab=16, depth=12)
    for ours, theirs in zip(model.blocks, ref_blocks):
        ours.load_state_dict(theirs.state_dict(), strict=True)
    model = model.to(dev).eval()
    with torch.no_grad():
        y = x.to(dev)
        for b in model.blocks:
            y = b(y)
    e = rel_l2(y.cpu().numpy(), h.numpy())
    assert e <= 1e-4, e


def test_long_context_two_pass_and_single_kernel_agree(fb, oracle, dev):
    """n_fft = 8192 / 16384: the two-pass path (radix-R streaming pass + 4096-point sub-transforms + streaming pass) and
    the single-kernel variants both match the oracle; ragged N, memory, bf16 included."""
    from fft_b200 import _lib
    lib = _lib.load()
    cases = [(2, 8192, 8192, 32, 16, True), (1, 16384, 16384, 48, 8, True), (2, 12000, 16384, 16, 16, False),
             (1, 20000, 16384, 16, 16, False)]
    for (B, N, n_fft, C, dg, with_mem) in cases:
        V, gate, mem = _rand_case(B, N, n_fft, C, dg, with_mem, seed=50 + N)
        want = oracle.mix_flat(V, gate, n_fft, dg, mem).numpy()
        args = (V.to(dev), gate.to(dev), None if mem is None else mem.to(dev))
        try:
            lib.spectre_mix_set_two_pass(2)                           # the three-launch path wherever it applies
            y2 = fb.spectral_mix(*args, n_fft=n_fft, group_width=dg)
            lib.spectre_mix_set_two_pass(0)                           # single-kernel variants only
            y1 = fb.spectral_mix(*args, n_fft=n_fft, group_width=dg)
            lib.spectre_mix_set_two_pass(1)                           # automatic (default)
            y0 = fb.spectral_mix(*args, n_fft=n_fft, group_width=dg)
        finally:
            lib.spectre_mix_set_two_pass(1)
        _check(y2, want)
        _check(y1, want)
        _check(y0, want)
    assert fb.plan_info(4, 16384, 16384, 768, 16)["launches"] == 3
    # n_fft = 8192 fp32: the TMEM-staged single kernel is the default (one pass over HBM, no workspace); bf16 keeps three launches
    info = fb.plan_info(4, 8192, 8192, 768, 16)
    assert info["launches"] == 1 and info["workspace_bytes"] == 0 and info["tile_channels"] == 4
    assert info["dit"] == 2 and info["radix"] == [16, 16, 16]          # even / odd rows, radix-2 combine in the middle pass
    assert fb.plan_info(4, 8191, 8192, 768, 16)["radix"] == [16, 2, 16, 16]   # odd row count: the plain single kernel
    assert fb.plan_info(4, 8192, 8192, 768, 16, torch.bfloat16)["launches"] == 3
    Vb = torch.randn(1, 16384, 32, generator=torch.Generator().manual_seed(60)).to(torch.bfloat16)
    gate = torch.randn(1, 2, 8193, dtype=torch.cfloat, generator=torch.Generator().manual_seed(61))
    y = fb.spectral_mix(Vb.to(dev), gate.to(dev), n_fft=16384, group_width=16)
    _check(y, oracle.mix_flat(Vb.float(), gate, 16384, 16).numpy(), rl2=REL_L2_BF16, mabs=2e-2)


def test_randomised_shapes(fb, oracle, dev):
    """Seeded random walk over the argument space: every supported n_fft, ragged N (both sides of n_fft), channel
    counts that leave partial tiles, group widths of every element mode, memory on/off, row-strided views."""
    rng = np.random.default_rng(1234)
    for trial in range(60):
        n_fft = int(2 ** rng.integers(5, 13))                     # 32 .. 4096
        N = int(rng.integers(1, int(n_fft * 1.3) + 2))
        dg = int(rng.choice([1, 2, 3, 4, 6, 8, 12, 16]))
        C = dg * int(rng.integers(1, 7))
        B = int(rng.integers(1, 4))
        with_mem = bool(rng.integers(0, 2))
        V, gate, mem = _rand_case(B, N, n_fft, C, dg, with_mem, seed=1000 + trial)
        want = oracle.mix_flat(V, gate, n_fft, dg, mem).numpy()
        Vd = V.to(dev)
        if trial % 3 == 0:                                         # a view with a row stride larger than C
            wide = torch.zeros(B, N, C + 8, device=dev)
            wide[:, :, 4:4 + C] = Vd
            Vd = wide[:, :, 4:4 + C]
        got = fb.spectral_mix(Vd, gate.to(dev), None if mem is None else mem.to(dev), n_fft=n_fft, group_width=dg)
        e1, e2 = rel_l2(got.cpu().numpy(), want), max_abs_rel(got.cpu().numpy(), want)
        assert e1 <= REL_L2_F32 and e2 <= MAX_ABS_F32, (trial, B, N, n_fft, C, dg, with_mem, e1, e2)


def test_cuda_graph_capture_and_second_stream(fb, dev):
    """The op is asynchronous on the current stream and holds no per-call host state: it can be captured into a CUDA graph
    after a warm-up call (twiddle tables / tensor maps are built on first use) and replayed, and it runs on a side stream."""
    V, gate, _ = _rand_case(3, 4096, 4096, 64, 16, False, seed=77)
    Vd, gd = V.to(dev), gate.to(dev)
    want = fb.spectral_mix(Vd, gd, n_fft=4096, group_width=16).clone()          # warm-up, eager result
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        y_side = fb.spectral_mix(Vd, gd, n_fft=4096, group_width=16)
    torch.cuda.current_stream().wait_stream(side)
    assert torch.equal(y_side, want)
    static_in = Vd.clone()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        static_out = fb.spectral_mix(static_in, gd, n_fft=4096, group_width=16)
    static_in.copy_(2.0 * Vd)                                                   # new data, same buffers
    graph.replay()
    torch.cuda.synchronize()
    ref2 = fb.spectral_mix(2.0 * Vd, gd, n_fft=4096, group_width=16)
    assert torch.equal(static_out, ref2)


@pytest.mark.parametrize("n_fft,B,C", [(256, 16, 64), (1024, 8, 768), (1024, 64, 768), (2048, 32, 768), (4096, 4, 768), (8192, 3, 768),
                                       (16384, 2, 768)])
def test_programmatic_dependent_launch_keeps_stream_order(n_fft, B, C, fb, dev):
    """Every mix launch carries the programmatic-stream-serialization attribute: its set-up may start while the previous kernel of
    the stream still runs, and griddepcontrol.wait orders all tensor accesses.  A chain of short dependent launches through the C
    ABI -- each reads what the previous one wrote (RAW) and overwrites what the previous one read (WAR) -- must give the bits
    of the same chain with the attribute switched off (sched bit 7), eagerly and as a captured CUDA graph."""
    import ctypes
    from fft_b200 import _lib
    lib = _lib.load()
    dg = 16
    gen = torch.Generator(device=dev).manual_seed(5 + n_fft)
    V0 = torch.randn(B, n_fft, C, device=dev, generator=gen)
    ang = torch.rand(B, C // dg, n_fft // 2 + 1, device=dev, generator=gen) * 6.2831853
    gate = torch.polar(torch.ones_like(ang), ang).contiguous()       # unit modulus: the chain keeps its scale
    gate[:, :, 0] = 1.0
    gate[:, :, -1] = -1.0
    ws_bytes = lib.spectre_mix_workspace_bytes(0, B, n_fft, n_fft, C, dg)
    ws = torch.empty(max(int(ws_bytes), 1), dtype=torch.uint8, device=dev)

    def chain(a, b, links=6):
        st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        for _ in range(links):
            rc = lib.spectre_mix_fwd_ws(a.data_ptr(), 0, a.stride(0), a.stride(1), gate.data_ptr(), None, C, b.data_ptr(), 0, b.stride(0),
                                        b.stride(1), B, n_fft, n_fft, C, dg, ws.data_ptr() if ws_bytes else None, int(ws_bytes), st)
            assert rc == 0, lib.spectre_mix_last_error().decode()
            a, b = b, a
        return a

    res = {}
    try:
        for name, flags in (("off", 3 | 128), ("on", 3)):
            lib.spectre_mix_set_sched(flags)
            a, b = V0.clone(), torch.zeros_like(V0)
            res[name] = chain(a, b).clone()
            a.copy_(V0)
            b.zero_()
            torch.cuda.synchronize()
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.stream(side):
                chain(a, b, links=2)                                  # warm-up on the capture stream
                a.copy_(V0)
                b.zero_()
                with torch.cuda.graph(graph, stream=side):
                    last = chain(a, b)
            graph.replay()
            torch.cuda.synchronize()
            res[name + "_graph"] = last.clone()
    finally:
        lib.spectre_mix_set_sched(3)
    assert torch.isfinite(res["off"]).all() and float(res["off"].abs().max()) > 0.1
    for k in ("on", "off_graph", "on_graph"):
        assert torch.equal(res[k], res["off"]), k


def test_tmem_variants_race_hunt():
    """Random batch / channel / row counts and dtypes at n_fft = 4096: the TMEM-staged variants under the default schedule
    (helper warpgroup phases, warp stagger, split barrier) agree bit for bit with the plain TMA variant."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "stress_4096.py"), "40", "7"], capture_output=True, text=True,
                       timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
