"""Shared test plumbing: the ``gpu`` marker, fixture loading, the C-oracle loader."""
import ctypes
import glob
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def golden_names(prefix):
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, prefix + "*.npz")))


def load_golden(name):
    with np.load(os.path.join(GOLDEN, name + ".npz")) as z:
        return {k: z[k] for k in z.files}


@pytest.fixture(scope="session")
def c_oracle():
    """The plain-C double-precision checker (oracle/spectre_mix_oracle.c), built on demand."""
    so = os.path.join(ROOT, "oracle", "libspectre_mix_oracle.so")
    src = os.path.join(ROOT, "oracle", "spectre_mix_oracle.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])
    lib = ctypes.CDLL(so)
    fp = ctypes.POINTER(ctypes.c_float)
    lib.spectre_mix_oracle_f64.argtypes = [fp, fp, fp, fp] + [ctypes.c_int] * 5
    lib.spectre_mix_oracle_f64.restype = ctypes.c_int

    def run(V, gate, mem, n_fft, group_width):
        V = np.ascontiguousarray(V, dtype=np.float32)
        gate = np.ascontiguousarray(gate, dtype=np.complex64)
        B, N, C = V.shape
        out = np.empty((B, min(N, n_fft), C), dtype=np.float32)
        memp = None
        if mem is not None:
            mem = np.ascontiguousarray(mem, dtype=np.complex64)
            memp = mem.ctypes.data_as(fp)
        rc = lib.spectre_mix_oracle_f64(V.ctypes.data_as(fp), gate.ctypes.data_as(fp), memp,
                                        out.ctypes.data_as(fp), B, N, n_fft, C, group_width)
        assert rc == 0, rc
        return out

    return run


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def max_abs_rel(a, b):
    """max |a-b| / max |b|  (the floor-free-of-zeros tolerance of SURVEY 8c)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))
