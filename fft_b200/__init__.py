"""fft_b200 -- B200-native Spectre spectral-mix forward path.

Drop-in for the hot path of jacobfa/fft's ``spectre.py`` (``SpectreHead.forward`` lines
506 and 542-553): batched real FFT along the sequence axis, complex multiply by the
per-sample / per-group gate (plus optional spectral memory), inverse real FFT.  The
arithmetic runs in ONE hand-written sm_100a kernel behind a C ABI
(``include/spectre_mix.h``); the rest of the block (projections, gate generator, norms,
MLP) stays stock PyTorch, exactly as in the reference.

There is no CPU path in this package: importing it works anywhere, but calling the op
without the CUDA library or a GPU raises.
"""
from .ops import gate_expand, rfft_seq, spectral_mix, spectral_mix_anchors, spectral_mix_host, plan_info  # noqa: F401
from .decode import PrefixFFTCache, decode_gate, head_decode_step  # noqa: F401
from .model import SpectreBase  # noqa: F401
from .modules import (  # noqa: F401
    ComplexModReLU,
    SpectreBlock,
    SpectreHead,
    SpectreMultiHead,
    WaveletRefinement,
    interp_complex_1d,
    patch_reference,
)

__all__ = [
    "spectral_mix", "spectral_mix_anchors", "spectral_mix_host", "rfft_seq", "gate_expand", "plan_info",
    "SpectreHead", "SpectreMultiHead", "SpectreBlock", "WaveletRefinement", "ComplexModReLU",
    "interp_complex_1d", "patch_reference", "PrefixFFTCache", "head_decode_step", "decode_gate", "SpectreBase",
]
