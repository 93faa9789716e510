"""Multi-GPU: shard the instance batch, one process per GPU, one gather.

Instances never interact, so the path shards with no
data-path collective: rank r owns the contiguous block `shard_range(B, W, r)`
of the batch axis and runs its own engine on its own GPU.  The only exchange
is the final gather of result tensors to rank 0 (`gather_rows`), a single
`torch.distributed` collective (NCCL over NVLink on the GPU box; gloo in the
CPU tests)."""
import torch
import torch.distributed as dist


def shard_range(B, world_size, rank):
    """Contiguous, balanced split: the first B % W ranks get one extra."""
    base, rem = divmod(int(B), int(world_size))
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi


def shard_sizes(B, world_size):
    return [shard_range(B, world_size, r)[1] - shard_range(B, world_size, r)[0]
            for r in range(world_size)]


def gather_rows(local, B, group=None, dst=0):
    """Gather the leading (instance) axis of `local` from every rank onto
    `dst`; returns the (B, ...) tensor on dst and None elsewhere.  Shards may
    be ragged by one row: they are padded to the largest shard so that a single
    equal-size all_gather moves the data."""
    if not (dist.is_available() and dist.is_initialized()) or \
            dist.get_world_size(group) == 1:
        return local
    W = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = shard_sizes(B, W)
    m = max(sizes)
    if local.shape[0] != sizes[rank]:
        raise ValueError("rank %d holds %d rows, expected %d" %
                         (rank, local.shape[0], sizes[rank]))
    pad = local
    if local.shape[0] < m:
        pad = torch.zeros((m,) + tuple(local.shape[1:]), dtype=local.dtype,
                          device=local.device)
        pad[:local.shape[0]] = local
    out = torch.empty((W * m,) + tuple(local.shape[1:]), dtype=local.dtype,
                      device=local.device)
    dist.all_gather_into_tensor(out, pad.contiguous(), group=group)
    if rank != dst:
        return None
    if all(s == m for s in sizes):
        return out
    parts = [out[r * m: r * m + sizes[r]] for r in range(W)]
    return torch.cat(parts, dim=0)


def run_sharded(circ, an_list, batch, group=None):
    """`ahkab_b200.run` on this rank's shard; raw result tensors gathered to
    rank 0.  Returns (results_of_this_shard, gathered_raw_or_None)."""
    from . import analyses
    W = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    lo, hi = shard_range(batch.B, W, rank)
    res = analyses.run(circ, an_list, batch=batch.slice(lo, hi))
    gathered = {}
    for an, sol in res.items():
        gathered[an] = {k: gather_rows(v, batch.B, group)
                        for k, v in sol.raw.items() if torch.is_tensor(v)}
    return res, (gathered if rank == 0 else None)
