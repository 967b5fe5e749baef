"""Batch sharding across GPUs.

Every (batch row, channel) column of the spectral mix is an independent transform, so the path shards over batch
rows with NO data-path collective: rank r of W owns a contiguous block of rows and runs the single-GPU kernel on
it.  ``torch.distributed`` (NCCL over NVLink on the GPU box, gloo in the CPU tests) only carries barriers and the
reductions of the measurement.
"""
from __future__ import annotations

from typing import Tuple

import torch


def shard_rows(total_rows: int, rank: int, world: int) -> Tuple[int, int]:
    """[begin, end) of the batch rows rank ``rank`` owns; blocks differ by at most one row and cover [0, total)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, rem = divmod(total_rows, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def micro_batches(rows: int, micro_batch: int):
    """[begin, end) spans that stream ``rows`` local rows through reused buffers in launches of at most ``micro_batch``
    rows (SURVEY 8e memory caveat: 206 GB of I/O tensors at batch 8192 do not fit one GPU)."""
    if micro_batch <= 0:
        raise ValueError(f"bad micro_batch {micro_batch}")
    return [(r0, min(r0 + micro_batch, rows)) for r0 in range(0, rows, micro_batch)]


def reduce_measurement(elapsed_ms: float, units: int, checksum: float, device=None):
    """(max elapsed over ranks, total units, summed checksum): the whole-job view of a sharded run.

    Works on any initialised process group (nccl or gloo); returns the inputs unchanged when not distributed.
    """
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return elapsed_ms, units, checksum
    t = torch.tensor([elapsed_ms], dtype=torch.float64, device=device)
    s = torch.tensor([float(units), float(checksum)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(s, op=dist.ReduceOp.SUM)
    return float(t.item()), int(round(float(s[0].item()))), float(s[1].item())
