"""CPU: the C-ABI library loads and exports every symbol include/spectre_mix.h declares.
No compute call is made (there is no GPU here)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "spectre_mix.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b(spectre_[a-z_0-9]+)\s*\(", src)
    return sorted(set(names))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    from fft_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        g.build()
    return ctypes.CDLL(_lib.LIB_PATH)


def test_header_symbols_are_exported(lib):
    from fft_b200 import _lib
    names = _declared_symbols()
    assert names, "no declarations parsed"
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/spectre_mix.h but not exported"
    assert sorted(_lib.SYMBOLS) == names


def test_abi_version_and_error_string(lib):
    lib.spectre_mix_abi_version.restype = ctypes.c_int
    assert lib.spectre_mix_abi_version() == 2
    lib.spectre_mix_last_error.restype = ctypes.c_char_p
    assert isinstance(lib.spectre_mix_last_error(), bytes)


def test_argument_validation_needs_no_gpu():
    """Bad arguments are rejected before any CUDA call (error codes of include/spectre_mix.h)."""
    from fft_b200 import _lib
    lib = _lib.load()
    # n_fft not a power of two -> UNSUPPORTED (2); C % group_width != 0 -> BAD_ARG (1)
    rc = lib.spectre_mix_fwd(None, 0, 0, 0, None, None, 0, None, 0, 0, 0, 1, 100, 100, 8, 4, None)
    assert rc == 2 and b"power of two" in lib.spectre_mix_last_error()
    rc = lib.spectre_mix_fwd(None, 0, 0, 0, None, None, 0, None, 0, 0, 0, 1, 128, 128, 10, 4, None)
    assert rc == 1
    # empty batch is a no-op success even without a device
    rc = lib.spectre_mix_fwd(None, 0, 0, 0, None, None, 0, None, 0, 0, 0, 0, 128, 128, 8, 4, None)
    assert rc == 0


def test_host_entry_chunk_schedule(monkeypatch):
    """Host logic of spectre_mix_fwd_host (no GPU): the batch is walked in chunks whose row counts ramp up by doubling from
    ~12 MB of V to ~100 MB and back down; every row is covered exactly once whatever the sizes."""
    import ctypes
    from fft_b200 import _lib
    lib = _lib.load()
    monkeypatch.delenv("SPECTRE_MIX_HOST_CHUNK_MB", raising=False)
    monkeypatch.delenv("SPECTRE_MIX_HOST_CHUNK_MAX_MB", raising=False)

    def sched(B, N, C):
        buf = (ctypes.c_int * 4096)()
        n = lib.spectre_mix_host_schedule(B, N, C, buf, 4096)
        assert 0 <= n <= 4096
        return list(buf[:n])

    s = sched(148, 4096, 768)                      # the bench's e2e step: 12.6 MB per row
    assert sum(s) == 148 and s[:3] == [1, 2, 4] and s[-3:] == [4, 2, 1] and max(s) == 8 and all(r >= 1 for r in s)
    mid = s[3:-3]
    assert all(r == 8 for r in mid[:-1]) and 1 <= mid[-1] <= 8
    assert sched(3, 4096, 768) == [1, 1, 1]        # too few rows for a ramp: the small chunk size throughout
    assert sched(0, 4096, 768) == []
    assert sched(5, 16384, 768) == [1, 2, 1, 1]        # 50 MB rows: the ramp is 1 -> 2 rows, the remainder chunk before the way down
    for (B, N, C) in [(1, 32, 4), (7, 100, 12), (600, 1024, 64), (8192, 4096, 768), (33, 20000, 8), (1000, 128, 64)]:
        s = sched(B, N, C)
        assert sum(s) == B and all(r >= 1 for r in s), (B, N, C, s)
        row = N * C * 4
        assert max(s) * row <= max(100 << 20, row)       # no chunk above ~100 MB unless a single row is
        half = len(s) // 2
        assert s[:half] == sorted(s[:half]) or len(set(s)) <= 2      # rising front ...
        assert s[-3:] == sorted(s[-3:], reverse=True) or len(s) < 3   # ... falling tail
    monkeypatch.setenv("SPECTRE_MIX_HOST_CHUNK_MB", "32")           # experiment knob: uniform chunks
    s = sched(148, 4096, 768)
    assert sum(s) == 148 and set(s[:-1]) == {2}
    assert lib.spectre_mix_host_schedule(-1, 4096, 768, None, 0) == -1


def test_no_cpu_fallback():
    """The product path must fail loudly on CPU tensors instead of computing on the host."""
    import torch
    import fft_b200
    V = torch.randn(1, 32, 4)
    gate = torch.ones(1, 1, 17, dtype=torch.cfloat)
    with pytest.raises(RuntimeError, match="no CPU"):
        fft_b200.spectral_mix(V, gate, n_fft=32, group_width=4)


def test_product_does_not_import_oracle():
    """oracle/ is test infrastructure: nothing under fft_b200/ may reference it."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "fft_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.replace("no CPU oracle", ""), f"{f} mentions oracle"
                assert "cufft" not in txt.lower().replace("no cufft", "").replace("not cufft", ""), f"{f} mentions cuFFT"


def test_bench_reference_arm_prints_contract_line():
    """`bench.py --impl reference` (CPU path, no GPU needed) prints one JSON line with the contract's keys."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["value"] > 0 and line["cpu_baseline"]["kind"] in ("port", "reference")
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
