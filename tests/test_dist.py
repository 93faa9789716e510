"""CPU: the N > 1 host logic (batch sharding + the single result gather) with two
gloo ranks on 127.0.0.1."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ahkab_b200 import dist as ahk_dist
from ahkab_b200 import workloads as W


def test_shard_range_partitions_the_batch():
    for B in (1, 7, 64, 65536, 16385):
        for Wn in (1, 2, 3, 8):
            parts = [ahk_dist.shard_range(B, Wn, r) for r in range(Wn)]
            assert parts[0][0] == 0 and parts[-1][1] == B
            assert all(parts[i][1] == parts[i + 1][0] for i in range(Wn - 1))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1
            assert sizes == ahk_dist.shard_sizes(B, Wn)


def test_param_batch_slice_matches_shards():
    w = W.c3_colpitts(101)
    full = w.batch()
    lo, hi = ahk_dist.shard_range(101, 2, 1)
    part = full.slice(lo, hi)
    assert part.B == hi - lo
    for key, v in full.items():
        assert np.array_equal(part[key], v[lo:hi])


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, B, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        lo, hi = ahk_dist.shard_range(B, world, rank)
        # stand-in for this rank's results: rows tagged with their global instance id
        x = torch.arange(lo, hi, dtype=torch.float64).reshape(-1, 1).repeat(1, 3)
        x[:, 1] += 0.5
        st = torch.full((hi - lo,), rank + 1, dtype=torch.int32)
        gx = ahk_dist.gather_rows(x, B)
        gs = ahk_dist.gather_rows(st, B)
        if rank == 0:
            q.put((gx.numpy(), gs.numpy()))
        else:
            assert gx is None and gs is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('B', [10, 11])
def test_gather_rows_two_gloo_ranks(B):
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, B, q)) for r in range(2)]
    for p in procs:
        p.start()
    gx, gs = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert gx.shape == (B, 3) and gs.shape == (B,)
    assert np.array_equal(gx[:, 0], np.arange(B))          # order = global instance order
    assert np.array_equal(gx[:, 1], np.arange(B) + 0.5)
    half = ahk_dist.shard_sizes(B, 2)[0]
    assert (gs[:half] == 1).all() and (gs[half:] == 2).all()


def test_gather_rows_single_process_is_identity():
    x = torch.arange(6.).reshape(3, 2)
    assert ahk_dist.gather_rows(x, 3) is x
