"""Spectre-base harness for BASELINE config 3 (the reference ships no full model).

``Embedding -> 12 x SpectreBlock(768, 12 heads, n_fft=4096) -> LayerNorm -> Linear`` assembled from the module shells;
everything except the spectral mix is stock PyTorch.  Under ``torch.autocast("cuda", torch.bfloat16)`` the projections
and MLPs run in bf16, the fused kernel takes bf16 V directly (bf16 in / fp32 math / bf16 out) and the gate generator
is kept in fp32 (``view_as_complex`` has no bf16 form -- the stock module cannot run this way).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .modules import SpectreBlock


class SpectreBase(nn.Module):
    def __init__(self, vocab: int = 32000, embed_dim: int = 768, num_heads: int = 12, n_fft: int = 4096, depth: int = 12,
                 **block_kw):
        super().__init__()
        kw = dict(mlp_ratio=4, d_gate=256, pooling_type="mean", num_groups=4, wavelet_on_rate=0.0, memory_size=0)
        kw.update(block_kw)
        self.embed = nn.Embedding(vocab, embed_dim)
        self.blocks = nn.ModuleList([SpectreBlock(embed_dim, num_heads, n_fft, **kw) for _ in range(depth)])
        self.norm = nn.LayerNorm(embed_dim)
        self.head = nn.Linear(embed_dim, vocab, bias=False)

    def forward(self, tokens: torch.Tensor, return_hidden: bool = False) -> torch.Tensor:
        x = self.embed(tokens)
        for blk in self.blocks:
            x = blk(x)
        x = self.norm(x)
        return x if return_hidden else self.head(x)
