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


def _worker_packed(rank, world, port, q):
    """ragged messages (what LTE-controlled transients produce) + a packed result set"""
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        n = 5 + 3 * rank                                   # ragged: 5 and 8 elements
        buf = torch.arange(n, dtype=torch.float64) + 100 * rank
        big, views = ahk_dist.gather_packed(buf)           # sizes exchanged (one small all_gather)
        # a packed result set like Engine._new_packed builds it: x [B][2] f64 | status [B] i32
        B = 3 + rank
        base = torch.zeros(B * 16 + ((B * 4 + 7) & ~7), dtype=torch.uint8)
        layout = [('x', (B, 2), torch.float64, 0, B * 16, B * 16),
                  ('status', (B,), torch.int32, B * 16, B * 4, (B * 4 + 7) & ~7)]
        loc = ahk_dist.unpack(base, layout)
        loc['x'][:] = rank + torch.arange(B * 2, dtype=torch.float64).reshape(B, 2) / 10
        loc['status'][:] = 7 + rank
        gb, gv = ahk_dist.gather_result({'_packed': (base, layout)}, ragged=True)
        if rank == 0:
            lay1 = [('x', (4, 2), torch.float64, 0, 64, 64), ('status', (4,), torch.int32, 64, 16, 16)]
            r1 = ahk_dist.unpack(gv[1], lay1)
            q.put((big.numpy(), [v.numpy().copy() for v in views], r1['x'].numpy().copy(),
                   r1['status'].numpy().copy()))
        else:
            assert big is None and views is None and gb is None
    finally:
        dist.destroy_process_group()


def test_gather_packed_ragged_two_gloo_ranks():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_packed, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    big, views, x1, st1 = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert np.array_equal(big, np.concatenate([np.arange(5.), np.arange(8.) + 100]))
    assert np.array_equal(views[0], np.arange(5.)) and np.array_equal(views[1], np.arange(8.) + 100)
    assert np.array_equal(x1, 1 + np.arange(8.).reshape(4, 2) / 10) and (st1 == 8).all()


def test_gather_rows_single_process_is_identity():
    x = torch.arange(6.).reshape(3, 2)
    assert ahk_dist.gather_rows(x, 3) is x
