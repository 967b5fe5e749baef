"""Pin the oracles (torch restatement + plain C) against the reference-generated goldens
and against analytic known-answer tests.  CPU only."""
import numpy as np
import pytest
import torch

from conftest import golden_names, load_golden, max_abs_rel, rel_l2
from oracle import spectre_mix_oracle as oracle

MIX_CASES = golden_names("mix_")


def _t(x):
    return None if x is None else torch.from_numpy(x)


@pytest.mark.parametrize("name", MIX_CASES)
def test_torch_oracle_matches_reference_goldens(name):
    g = load_golden(name)
    n_fft, H, dg = int(g["n_fft"]), int(g["num_heads"]), int(g["group_width"])
    mem = _t(g.get("mem"))
    y_loop = oracle.mix_head_loop(_t(g["V"]), _t(g["gate"]), n_fft, H, mem).numpy()
    y_flat = oracle.mix_flat(_t(g["V"]), _t(g["gate"]), n_fft, dg, mem).numpy()
    # same library, same call sites: the head loop must reproduce the reference bit for bit
    assert np.array_equal(y_loop, g["out"])
    assert rel_l2(y_flat, g["out"]) < 1e-6 and max_abs_rel(y_flat, g["out"]) < 1e-5


@pytest.mark.parametrize("name", MIX_CASES)
def test_c_oracle_matches_reference_goldens(name, c_oracle):
    g = load_golden(name)
    y = c_oracle(g["V"], g["gate"], g.get("mem"), int(g["n_fft"]), int(g["group_width"]))
    assert y.shape == g["out"].shape
    # C oracle computes in double; the golden is the reference's fp32 (MKL) result
    assert rel_l2(y, g["out"]) < 2e-6 and max_abs_rel(y, g["out"]) < 1e-5


def test_c_oracle_matches_fp64_torch(c_oracle):
    torch.manual_seed(0)
    V = torch.randn(2, 50, 12)
    gate = torch.randn(2, 3, 33, dtype=torch.cfloat)
    mem = torch.randn(33, 12, dtype=torch.cfloat)
    y64 = oracle.mix_flat(V, gate, 64, 4, mem, dtype=torch.float64).numpy()
    y = c_oracle(V.numpy(), gate.numpy(), mem.numpy(), 64, 4)
    assert rel_l2(y, y64) < 2e-7
    # non power of two transform length goes through the O(n^2) definition
    gate2 = torch.randn(2, 3, 31, dtype=torch.cfloat)
    y64 = oracle.mix_flat(V, gate2, 60, 4, None, dtype=torch.float64).numpy()
    y = c_oracle(V.numpy(), gate2.numpy(), None, 60, 4)
    assert rel_l2(y, y64) < 2e-7


def test_truncation_when_N_exceeds_n_fft(c_oracle):
    # spectre.py:506 truncates to n_fft rows and the output shrinks
    torch.manual_seed(1)
    V = torch.randn(1, 40, 4)
    gate = torch.randn(1, 1, 17, dtype=torch.cfloat)
    y = oracle.mix_flat(V, gate, 32, 4)
    assert y.shape == (1, 32, 4)
    yc = c_oracle(V.numpy(), gate.numpy(), None, 32, 4)
    assert rel_l2(yc, y.numpy()) < 2e-6


# ---------------- analytic known-answer tests (no reference needed) ----------------
@pytest.mark.parametrize("n_fft,N", [(64, 64), (64, 40), (256, 256)])
def test_kat_identity_gate(n_fft, N, c_oracle):
    torch.manual_seed(2)
    V = torch.randn(2, N, 8)
    gate = torch.ones(2, 2, n_fft // 2 + 1, dtype=torch.cfloat)
    for y in (oracle.mix_flat(V, gate, n_fft, 4).numpy(), c_oracle(V.numpy(), gate.numpy(), None, n_fft, 4)):
        assert max_abs_rel(y, V.numpy()) < 1e-5


def test_kat_shift_gate(c_oracle):
    # gate[k] = exp(-2 pi i k s / n) is a circular shift by s (linear when N + s <= n_fft)
    n_fft, s, N = 128, 5, 100
    k = torch.arange(n_fft // 2 + 1)
    gate = torch.exp(-2j * torch.pi * k * s / n_fft).to(torch.cfloat).expand(1, 1, -1).contiguous()
    torch.manual_seed(3)
    V = torch.randn(1, N, 4)
    want = torch.zeros(1, N, 4)
    want[:, s:] = V[:, : N - s]
    for y in (oracle.mix_flat(V, gate, n_fft, 4).numpy(), c_oracle(V.numpy(), gate.numpy(), None, n_fft, 4)):
        assert np.abs(y - want.numpy()).max() < 1e-5


def test_kat_dc_nyquist_imag_ignored_and_memory_only():
    torch.manual_seed(4)
    n_fft = 64
    V = torch.randn(2, 64, 8)
    gate = torch.randn(2, 2, 33, dtype=torch.cfloat)
    mem = torch.randn(33, 8, dtype=torch.cfloat)
    y0 = oracle.mix_flat(V, gate, n_fft, 4, mem)
    # zero gate: every batch row is irfft(memory)
    yz = oracle.mix_flat(V, torch.zeros_like(gate), n_fft, 4, mem)
    want = torch.fft.irfft(mem, n=n_fft, dim=0)
    assert torch.allclose(yz[0], want, atol=1e-6) and torch.allclose(yz[1], want, atol=1e-6)
    # imag of memory's bin 0 and bin n/2 has no effect
    mem2 = mem.clone()
    mem2[0] = torch.complex(mem[0].real, torch.randn(8))
    mem2[-1] = torch.complex(mem[-1].real, torch.randn(8))
    assert torch.equal(oracle.mix_flat(V, gate, n_fft, 4, mem2), y0)


def test_decode_restatements_match_reference_goldens():
    g = load_golden("decode_prefill_n256_d32")
    spec = oracle.prefill_spectrum(_t(g["V"]), 256).numpy()
    assert np.array_equal(spec, g["prefix_fft"])
    for p, want in zip(g["pruned_pos"], g["pruned"]):
        got = oracle.pruned_irfft_single(_t(g["X_half"]), 256, int(p)).numpy()
        assert np.allclose(got, want, rtol=1e-5, atol=1e-6)


def test_cache_update_restatement_matches_reference_sequence():
    """Replay the reference's 100 decode steps (with eviction) through the oracle's spectrum update."""
    g = load_golden("decode_seq_n256_d32")
    n = 256
    Vp, vs = _t(g["Vp"]), _t(g["vs"])
    prefix = oracle.prefill_spectrum(Vp, n)
    V_buf = torch.zeros(n, Vp.shape[1])
    V_buf[: Vp.shape[0]] = Vp
    t = Vp.shape[0] - 1
    for i in range(vs.shape[0]):
        t += 1
        oracle.cache_update(prefix, V_buf, vs[i], t, n)
    assert t == int(g["t"])
    assert np.array_equal(prefix.numpy(), g["prefix_fft"])


# --------------------------------------------------------------------------- gate generator tail
GATE_CASES = golden_names("gate_")


@pytest.mark.parametrize("name", GATE_CASES)
def test_gate_tail_oracle_matches_reference_goldens(name):
    g = load_golden(name)
    F_half = int(g["n_fft"]) // 2 + 1
    a, bias, eps = _t(g["anchors"]), _t(g["bias"]), _t(g["eps"])
    # same library call sites as spectre.py:41-60, :110-121, :536 -> bit for bit
    assert np.array_equal(oracle.gate_tail(a, bias, eps, None, F_half).numpy(), g["gate"])
    assert np.array_equal(oracle.gate_tail(a, bias, eps, _t(g["pos1"]), F_half).numpy(), g["gate_pos1"])
    assert np.array_equal(oracle.gate_tail(a, bias, eps, _t(g["posB"]), F_half).numpy(), g["gate_posB"])


def _crel(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / np.linalg.norm(np.asarray(b)))


@pytest.mark.parametrize("name", GATE_CASES)
def test_gate_tail_direct_restatement_matches_reference_goldens(name):
    """Keys cubic convolution (A = -0.75, border clamp, align_corners) written out == what grid_sample evaluates."""
    g = load_golden(name)
    F_half = int(g["n_fft"]) // 2 + 1
    a = g["anchors"]
    up = oracle.gate_tail_direct(a, np.zeros_like(g["bias"]), 0.0, None, F_half)   # scale = |z| / |z| = 1 -> pure interpolation
    assert _crel(up, g["interp"]) < 3e-6      # the reference's float32 sample positions cost ~1e-6 against exact ones
    for key, pos in (("gate", None), ("gate_pos1", g["pos1"]), ("gate_posB", g["posB"])):
        y = oracle.gate_tail_direct(a, g["bias"], g["eps"], pos, F_half)
        assert _crel(y, g[key]) < 3e-6 and np.abs(y - g[key]).max() < 1e-5 * np.abs(g[key]).max()


def test_batched_gate_generator_equals_per_head_loop(monkeypatch):
    """Host logic of the all-heads gate generator (pool -> LayerNorm -> MLP as batched GEMMs over stacked head weights)
    against the per-head generator, on CPU with the expansion kernel replaced by its stock-PyTorch statement."""
    import fft_b200.modules as M
    from fft_b200 import ops
    monkeypatch.setattr(M, "gate_expand",
                        lambda a, b, e, p=None, *, F_half, G: ops._gate_expand_torch(a, b, e, p, F_half, G))
    torch.manual_seed(3)
    mh = M.SpectreMultiHead(48, 3, 128, pooling_type="mean", wavelet_on_rate=0.0, num_groups=4).eval()
    with torch.no_grad():
        for h in mh.heads:
            h.modrelu.bias.add_(0.3 * torch.randn_like(h.modrelu.bias))
    assert M._heads_batchable(mh)
    x = torch.randn(2, 100, 48)
    pos = torch.polar(torch.ones(2, 65), torch.rand(2, 65) * 6.28)
    with torch.no_grad():
        Q_all = torch.stack([h.W_q(c) for h, c in zip(mh.heads, torch.chunk(x, 3, -1))], dim=2)
        for p in (pos, pos[:1], None):
            gate, q_pool = M.multihead_gate(mh, Q_all, p)
            per = [M.head_gate(h, Q_all[:, :, i], p) for i, h in enumerate(mh.heads)]
            assert _crel(gate.numpy(), torch.cat([g for g, _ in per], 1).numpy()) < 1e-5
            assert torch.allclose(q_pool, torch.cat([q for _, q in per], -1), atol=1e-5)
    # cached stacks follow in-place parameter updates
    with torch.no_grad():
        mh.heads[1].gate_mlp[2].bias.add_(1.0)
        gate2, _ = M.multihead_gate(mh, Q_all, None)
        per2 = torch.cat([M.head_gate(h, Q_all[:, :, i], None)[0] for i, h in enumerate(mh.heads)], 1)
    assert _crel(gate2.numpy(), per2.numpy()) < 1e-5 and _crel(gate2.numpy(), gate.numpy()) > 1e-3
    # ... and writes through .data (EMA swaps, old-style optimizers: no version bump, same storage) are seen too
    with torch.no_grad():
        mh.heads[0].gate_mlp[2].bias.data.copy_(mh.heads[0].gate_mlp[2].bias.data + 2.0)
        gate3, _ = M.multihead_gate(mh, Q_all, None)
        per3 = torch.cat([M.head_gate(h, Q_all[:, :, i], None)[0] for i, h in enumerate(mh.heads)], 1)
    assert _crel(gate3.numpy(), per3.numpy()) < 1e-5 and _crel(gate3.numpy(), gate2.numpy()) > 1e-3
