"""world_size-2 gloo test of the N>1 path: row sharding, barrier, MAX/SUM reductions (CPU only).

The shards are transformed with the CPU oracle here (the CUDA kernel needs a GPU); what is under test is the
host logic bench.py --gpus N relies on: the shards tile the batch exactly, need no exchange, and the reduced
measurement is the whole-job view.
"""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fft_b200.dist import micro_batches, reduce_measurement, shard_rows


def test_shard_rows_cover_batch_exactly():
    for total in (0, 1, 7, 8, 8192, 8193):
        for world in (1, 2, 3, 8):
            spans = [shard_rows(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1


def test_cfg5_strong_scaling_plan():
    """BASELINE.json configs[4]: 8192 rows over 1/2/4/8 ranks, each rank streaming micro-batches of <= 148 rows: every row of
    the global batch is processed exactly once per step at every world size."""
    for world in (1, 2, 4, 8):
        seen = []
        for r in range(world):
            b0, b1 = shard_rows(8192, r, world)
            spans = micro_batches(b1 - b0, 148)
            assert all(0 < e - b <= 148 for b, e in spans)
            assert spans[0][0] == 0 and spans[-1][1] == b1 - b0 and all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            seen += [b0 + i for b, e in spans for i in range(b, e)]
        assert seen == list(range(8192))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import spectre_mix_oracle as oracle
    torch.manual_seed(0)                      # every rank builds the same global problem, then keeps its rows
    B, N, C, dg = 5, 64, 16, 4
    V = torch.randn(B, N, C)
    gate = torch.randn(B, C // dg, N // 2 + 1, dtype=torch.cfloat)
    b0, b1 = shard_rows(B, rank, world)
    y_local = oracle.mix_flat(V[b0:b1], gate[b0:b1], N, dg)       # no data from other ranks is needed
    dist.barrier()
    elapsed, units, checksum = reduce_measurement(10.0 + rank, (b1 - b0) * N, float(y_local.double().sum()))
    y_full = oracle.mix_flat(V, gate, N, dg)
    ok = (abs(elapsed - (10.0 + world - 1)) < 1e-12 and units == B * N
          and abs(checksum - float(y_full.double().sum())) < 1e-6
          and torch.equal(y_local, y_full[b0:b1]))
    out[rank] = ok
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_batch_shard_gloo():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert dict(out) == {0: True, 1: True}
