"""Object sharding across the GPUs of one box + all-gather of the feature table.

Objects are independent (the reference's loop body touches one image at a time, NB:358-364), so
rank r owns the contiguous object range [r*ceil(N/G), min(N, (r+1)*ceil(N/G))); every rank pads
its block to ceil(N/G) rows so that one ``all_gather_into_tensor`` assembles the table.  No other
collective is on the data path.  Works with NCCL (cuda tensors) and gloo (cpu tensors, used by the
CPU tests of this logic).
"""
import math


def shard_range(n_objects, world_size, rank):
    """Contiguous shard [start, stop) of rank ``rank`` and the padded rows per rank."""
    per = int(math.ceil(n_objects / float(world_size))) if n_objects else 0
    start = min(n_objects, rank * per)
    stop = min(n_objects, start + per)
    return start, stop, per


def balanced_ranges(weights, world_size):
    """Contiguous ranges balanced by total weight (e.g. h_i*w_i for variable-size objects).
    Returns a list of (start, stop) per rank; ranges are contiguous and cover [0, N).  Boundary r
    is the prefix position whose cumulative weight is nearest to total*r/world_size."""
    n = len(weights)
    prefix = [0.0]
    for wgt in weights:
        prefix.append(prefix[-1] + float(wgt))
    total = prefix[-1]
    bounds, pos = [0], 0
    for r in range(1, world_size):
        target = total * r / world_size
        while pos < n and abs(prefix[pos + 1] - target) <= abs(prefix[pos] - target):
            pos += 1
        bounds.append(pos)
    bounds.append(n)
    return [(bounds[k], bounds[k + 1]) for k in range(world_size)]


def gather_table(local_rows, n_objects, group=None, out=None, chunk_rows=None, side_stream=None):
    """All-gather per-rank row blocks into the full [n_objects, F] table on every rank.

    local_rows : [rows_r, F] tensor holding this rank's shard (rows_r <= per)
    chunk_rows : optional; gather in slabs of this many rows (lets the caller overlap slab k's
                 gather with slab k+1's kernels by issuing it on ``side_stream``)
    """
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    start, stop, per = shard_range(n_objects, world, rank)
    F = local_rows.shape[1]
    assert local_rows.shape[0] == stop - start
    if per == 0:
        return local_rows.new_empty((0, F))
    padded = local_rows
    if stop - start < per:
        padded = local_rows.new_zeros((per, F))
        padded[: stop - start] = local_rows
    padded = padded.contiguous()
    full = out if out is not None else local_rows.new_empty((world * per, F))
    if chunk_rows is None or chunk_rows >= per:
        dist.all_gather_into_tensor(full, padded, group=group)
    else:
        # slab-wise: slab s of every rank lands in a [world, rows, F] staging block
        for s0 in range(0, per, chunk_rows):
            s1 = min(per, s0 + chunk_rows)
            stage = local_rows.new_empty((world, s1 - s0, F))
            dist.all_gather_into_tensor(stage.view(-1, F), padded[s0:s1].contiguous(), group=group)
            full.view(world, per, F)[:, s0:s1].copy_(stage)
    return full[:n_objects] if world * per != n_objects else full


def extract_sharded(extractor, make_shard, n_objects, group=None):
    """Run ``extractor`` on this rank's shard and gather.  ``make_shard(start, stop)`` returns
    the keyword arguments of ``FeatureExtractor.extract_planar`` for objects [start, stop)."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    start, stop, _ = shard_range(n_objects, world, rank)
    local = extractor.extract_planar(**make_shard(start, stop))
    return gather_table(local, n_objects, group=group)
