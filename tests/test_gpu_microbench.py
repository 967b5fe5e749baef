"""The tensor-memory transpose behind DESIGN 3.9 (``tools/microbench/tmem_xchg.cu``): compiled and run on the GPU box.

Not part of the product path -- it pins the claim that a ``tcgen05.st.32x32b`` + ``tcgen05.ld.16x256b`` pair (twice) is a 16 x 16
transpose among the lanes of a half-warp and that the shape-swapped pair is its inverse, which the optional
``-DSPX_TMEMX`` build of the mix kernel relies on.
"""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_tmem_transpose_exchange(tmp_path):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not on this box")
    exe = tmp_path / "tmem_xchg"
    subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-o", str(exe),
                    os.path.join(ROOT, "tools", "microbench", "tmem_xchg.cu")], check=True, timeout=300)
    out = subprocess.run([str(exe)], check=True, timeout=120, capture_output=True, text=True).stdout
    assert "check: 0 mismatches" in out, out
    # the exchange through tensor memory must not be slower than the one through shared memory (measured 1400 vs 2050 cycles)
    cyc = {}
    for line in out.splitlines():
        if line.startswith("mode "):
            mode = int(line.split()[1])
            cyc[mode] = float(line.rsplit("=", 1)[1].split()[0])
    assert cyc[1] < cyc[2], out
