"""ctypes binding of the C ABI declared in ``include/spectre_mix.h``.

The library is built in-tree (``fft_b200/_C/libspectre_mix.so``) by
``__graft_entry__.build()`` / ``make -C fft_b200/csrc``.  Loading is lazy and LOUD: a
missing library raises ``RuntimeError`` -- there is no fallback path.
"""
from __future__ import annotations

import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_C", "libspectre_mix.so")

F32, BF16 = 0, 1

# every symbol include/spectre_mix.h declares (checked by tests/test_abi.py)
SYMBOLS = (
    "spectre_mix_abi_version",
    "spectre_mix_last_error",
    "spectre_mix_fwd",
    "spectre_mix_fwd_ws",
    "spectre_mix_host_release",
    "spectre_mix_host_schedule",
    "spectre_mix_fwd_anchors",
    "spectre_mix_anchors_workspace_bytes",
    "spectre_mix_workspace_bytes",
    "spectre_mix_dgate",
    "spectre_mix_fwd_host",
    "spectre_rfft_fwd",
    "spectre_decode_update",
    "spectre_decode_readout",
    "spectre_decode_step",
    "spectre_decode_workspace_bytes",
    "spectre_decode_gate",
    "spectre_gate_expand",
    "spectre_mix_plan",
    "spectre_mix_set_tile_channels",
    "spectre_mix_set_prefetch",
    "spectre_mix_set_tma",
    "spectre_mix_set_timeline",
    "spectre_mix_set_tmem",
    "spectre_mix_set_skew_ns",
    "spectre_mix_set_sched",
    "spectre_mix_set_l2_promotion",
    "spectre_mix_set_two_pass",
)


class PlanInfo(ctypes.Structure):
    _fields_ = [
        ("n_fft", ctypes.c_int),
        ("radix", ctypes.c_int * 4),
        ("tile_channels", ctypes.c_int),
        ("threads", ctypes.c_int),
        ("ctas_per_sm", ctypes.c_int),
        ("smem_bytes", ctypes.c_int),
        ("grid", ctypes.c_int),
        ("launches", ctypes.c_int),
        ("algorithmic_bytes", ctypes.c_int64),
        ("workspace_bytes", ctypes.c_int64),
        ("dit", ctypes.c_int),
    ]


_lib = None
_lock = threading.Lock()


def load():
    """Return the loaded library; raise if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"fft_b200: CUDA library not built: {LIB_PATH} is missing. "
                "Run `python -c 'import __graft_entry__ as g; g.build()'` or `make -C fft_b200/csrc`. "
                "There is no CPU fallback for the spectral-mix path."
            )
        lib = ctypes.CDLL(LIB_PATH)
        vp, i32, i64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64
        lib.spectre_mix_abi_version.restype = i32
        lib.spectre_mix_abi_version.argtypes = []
        lib.spectre_mix_last_error.restype = ctypes.c_char_p
        lib.spectre_mix_last_error.argtypes = []
        lib.spectre_mix_fwd.restype = i32
        lib.spectre_mix_fwd.argtypes = [vp, i32, i64, i64, vp, vp, i64, vp, i32, i64, i64, i32, i32, i32, i32, i32, vp]
        lib.spectre_mix_fwd_ws.restype = i32
        lib.spectre_mix_fwd_ws.argtypes = [vp, i32, i64, i64, vp, vp, i64, vp, i32, i64, i64, i32, i32, i32, i32, i32, vp,
                                           ctypes.c_size_t, vp]
        lib.spectre_mix_workspace_bytes.restype = ctypes.c_size_t
        lib.spectre_mix_workspace_bytes.argtypes = [i32, i32, i32, i32, i32, i32]
        lib.spectre_mix_fwd_anchors.restype = i32
        lib.spectre_mix_fwd_anchors.argtypes = [vp, i32, i64, i64, vp, vp, vp, vp, i64, i32, i32, vp, i64, vp, i32, i64, i64,
                                                i32, i32, i32, i32, i32, vp, ctypes.c_size_t, vp]
        lib.spectre_mix_anchors_workspace_bytes.restype = ctypes.c_size_t
        lib.spectre_mix_anchors_workspace_bytes.argtypes = [i32, i32, i32, i32, i32, i32]
        lib.spectre_mix_dgate.restype = i32
        lib.spectre_mix_dgate.argtypes = [vp, vp, i32, i64, i64, i64, i64, vp, i32, i32, i32, i32, i32, vp]
        lib.spectre_mix_fwd_host.restype = i32
        lib.spectre_mix_fwd_host.argtypes = [vp, vp, vp, vp, i32, i32, i32, i32, i32]
        lib.spectre_mix_host_release.restype = i32
        lib.spectre_mix_host_release.argtypes = []
        lib.spectre_mix_host_schedule.restype = i32
        lib.spectre_mix_host_schedule.argtypes = [i32, i32, i32, ctypes.POINTER(ctypes.c_int), i32]
        lib.spectre_rfft_fwd.restype = i32
        lib.spectre_rfft_fwd.argtypes = [vp, i32, i64, i64, vp, i32, i32, i32, i32, vp]
        lib.spectre_decode_update.restype = i32
        lib.spectre_decode_update.argtypes = [vp, vp, vp, i32, i32, ctypes.c_longlong, vp]
        lib.spectre_decode_readout.restype = i32
        lib.spectre_decode_readout.argtypes = [vp, vp, vp, i32, i32, i32, i32, vp, ctypes.c_size_t, vp]
        lib.spectre_decode_step.restype = i32
        lib.spectre_decode_step.argtypes = [vp, vp, vp, vp, vp, i32, i32, i32, ctypes.c_longlong, vp, ctypes.c_size_t, vp]
        lib.spectre_decode_workspace_bytes.restype = ctypes.c_size_t
        lib.spectre_decode_workspace_bytes.argtypes = [i32, i32]
        lib.spectre_decode_gate.restype = i32
        lib.spectre_decode_gate.argtypes = [vp, vp, vp, vp, i32, i32, i32, i32, ctypes.c_longlong, i32, vp]
        lib.spectre_gate_expand.restype = i32
        lib.spectre_gate_expand.argtypes = [vp, vp, vp, vp, ctypes.c_longlong, vp, i32, i32, i32, i32, i32, vp]
        lib.spectre_mix_plan.restype = i32
        lib.spectre_mix_plan.argtypes = [i32, i32, i32, i32, i32, i32, i32, i32, ctypes.POINTER(PlanInfo)]
        lib.spectre_mix_set_tile_channels.restype = i32
        lib.spectre_mix_set_tile_channels.argtypes = [i32]
        lib.spectre_mix_set_prefetch.restype = i32
        lib.spectre_mix_set_prefetch.argtypes = [i32]
        lib.spectre_mix_set_tma.restype = i32
        lib.spectre_mix_set_tma.argtypes = [i32]
        lib.spectre_mix_set_two_pass.restype = i32
        lib.spectre_mix_set_two_pass.argtypes = [i32]
        lib.spectre_mix_set_skew_ns.restype = i32
        lib.spectre_mix_set_skew_ns.argtypes = [i32]
        lib.spectre_mix_set_l2_promotion.restype = i32
        lib.spectre_mix_set_l2_promotion.argtypes = [i32]
        lib.spectre_mix_set_sched.restype = i32
        lib.spectre_mix_set_sched.argtypes = [i32]
        lib.spectre_mix_set_tmem.restype = i32
        lib.spectre_mix_set_tmem.argtypes = [i32]
        lib.spectre_mix_set_timeline.restype = i32
        lib.spectre_mix_set_timeline.argtypes = [vp]
        _lib = lib
    return _lib


def check(rc: int, what: str):
    if rc != 0:
        msg = load().spectre_mix_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"fft_b200.{what} failed (code {rc}): {msg}")
