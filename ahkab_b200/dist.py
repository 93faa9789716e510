"""Multi-GPU: shard the instance batch, one process per GPU, ONE result gather.

Instances never interact, so the path shards with no data-path
collective: rank r owns the contiguous block `shard_range(B, W, r)` of the batch
axis and runs its own engine on its own GPU.  The only exchange is the gather
of the results onto rank 0:

* every analysis call returns its result tensors as views into ONE device
  buffer (`raw['_packed']`, engine.py), so a rank's whole result set is one
  contiguous message;
* `gather_packed` moves those messages to rank 0 in ONE grouped point-to-point
  exchange (`batch_isend_irecv`: NCCL coalesces it into a single send/recv group
  over NVLink / NVSwitch; gloo in the CPU tests) — a gather, not an all-gather:
  only rank 0 pays for the memory and the traffic;
* ragged results (LTE-controlled transients: the packed rows differ per rank)
  are sized by one small all_gather of the byte counts first.
"""
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


def _active(group):
    return dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1


def gather_packed(buf, sizes=None, group=None, dst=0, out=None):
    """Gather one contiguous 1-D tensor per rank onto `dst`.

    buf: this rank's message (any dtype, 1-D, contiguous).  sizes: element counts of
    every rank's message if the caller knows them (fixed-size results); None = ragged,
    exchanged with one small all_gather.  Returns (big, views) on dst — `big` the
    concatenation in rank order, `views[r]` rank r's part of it — and (None, None)
    elsewhere.  `out`: optional preallocated destination on dst (>= sum(sizes))."""
    if not _active(group):
        return buf, [buf]
    W = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if sizes is None:
        mine = torch.tensor([buf.numel()], dtype=torch.int64, device=buf.device)
        allsz = torch.empty(W, dtype=torch.int64, device=buf.device)
        dist.all_gather_into_tensor(allsz, mine, group=group)
        sizes = [int(v) for v in allsz.tolist()]
    if int(sizes[rank]) != buf.numel():
        raise ValueError("rank %d sends %d elements, sizes says %d" % (rank, buf.numel(), sizes[rank]))
    if rank != dst:
        if buf.numel():
            for req in dist.batch_isend_irecv([dist.P2POp(dist.isend, buf, dst, group=group)]):
                req.wait()
        return None, None
    total = int(sum(sizes))
    big = out if out is not None and out.numel() >= total and out.dtype == buf.dtype else \
        torch.empty(total, dtype=buf.dtype, device=buf.device)
    views, ops, off = [], [], 0
    for r in range(W):
        v = big[off:off + int(sizes[r])]
        views.append(v)
        off += int(sizes[r])
        if r == rank:
            v.copy_(buf)
        elif v.numel():
            ops.append(dist.P2POp(dist.irecv, v, r, group=group))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    return big[:total], views


def gather_rows(local, B, group=None, dst=0):
    """Gather the leading (instance) axis of `local` from every rank onto `dst`;
    returns the (B, ...) tensor on dst and None elsewhere (shards may differ by one
    row).  One grouped send/recv, rank 0 only holds the result."""
    if not _active(group):
        return local
    W = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = shard_sizes(B, W)
    if local.shape[0] != sizes[rank]:
        raise ValueError("rank %d holds %d rows, expected %d" %
                         (rank, local.shape[0], sizes[rank]))
    per_row = 1
    for d in local.shape[1:]:
        per_row *= int(d)
    big, _ = gather_packed(local.contiguous().reshape(-1), [s * per_row for s in sizes],
                           group=group, dst=dst)
    if rank != dst:
        return None
    return big.reshape((B,) + tuple(local.shape[1:]))


def unpack(view, layout):
    """Result tensors of one rank from its part of a gathered buffer (`layout` = the
    items of raw['_packed'], see Engine._new_packed)."""
    out = {}
    for k, shape, dtype, off, nbytes, _nb in layout:
        out[k] = view[off:off + nbytes].view(dtype).view(shape)
    return out


def gather_result(raw, group=None, dst=0, ragged=False, out=None):
    """The single result collective of a sharded analysis: every rank's packed result
    buffer (raw['_packed']) to rank `dst`.  Returns (big, views) like gather_packed; use
    `unpack(views[r], layout_r)` to get rank r's tensors back.  ragged=False asserts that
    every rank's buffer has this rank's size (equal shards of a fixed-size result)."""
    base, _layout = raw['_packed']
    sizes = None
    if not ragged and _active(group):
        sizes = [base.numel()] * dist.get_world_size(group)
    return gather_packed(base, sizes, group=group, dst=dst, out=out)


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
