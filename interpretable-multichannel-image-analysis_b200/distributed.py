"""Object sharding across the GPUs of one box and assembly of the feature table on every rank.

Objects are independent (the reference's loop body touches one image at a time, NB:358-364), so
rank r owns the contiguous object range [r*ceil(N/G), min(N, (r+1)*ceil(N/G))); the only exchange on
the path is the gather of the per-rank row blocks (one all-gather of float64[N/G, F]).

``extract_sharded`` is the product call: it works the rank's shard off in slabs; the kernels write a
slab's rows straight into this rank's copy of the full table, and while the kernels of slab k+1 run,
slab k travels to the other ranks on a side stream.  Two transports:

  "p2p"         the tables are one symmetric-memory allocation (torch.distributed._symmetric_memory: every rank
                has the other ranks' tables mapped into its own address space); a rank pushes its rows with
                plain device-to-device copies: copy engines over NVLink, no SM is taken from the kernels
                (an NCCL all-gather is an SM-resident kernel that competes with the persistent grids).
  "collective"  ``all_gather_into_tensor`` per slab (NCCL, or gloo on CPU tensors) -- the contract the
                p2p transport is validated against, and the path the CPU tests exercise.
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
    """All-gather per-rank row blocks into the full [n_objects, F] table on every rank (blocking).

    local_rows : [rows_r, F] tensor holding this rank's shard (rows_r <= per)
    chunk_rows : optional; gather in slabs of this many rows
    """
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


class ShardedTable:
    """This rank's copy of the full table ([world * per, F], rank r's rows at r * per) and the means to fill
    the other ranks' copies with this rank's rows, slab by slab, off the compute stream."""

    def __init__(self, n_objects, width, device=None, group=None, transport=None, dtype=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.group = torch, dist, group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.n_objects, self.width = int(n_objects), int(width)
        self.start, self.stop, self.per = shard_range(self.n_objects, self.world, self.rank)
        dtype = torch.float64 if dtype is None else dtype
        self.device = torch.device("cpu") if device is None else torch.device(device)
        self.cuda = self.device.type == "cuda"
        if transport is None:
            transport = "p2p" if (self.cuda and self.world > 1) else "collective"
        self.transport = transport
        self.side = torch.cuda.Stream(device=self.device) if self.cuda else None
        self.peers = None
        self.full = None
        self._stage = {}
        self._n_push = 0
        shape = (self.world * self.per, self.width)
        if self.transport == "p2p":
            try:
                self._map_peers(shape, dtype)
            except Exception as exc:                       # no symmetric memory on this box: the collective still works
                self.transport, self.fallback_reason, self.peers, self.full = "collective", repr(exc), None, None
        if self.full is None:
            self.full = torch.zeros(shape, dtype=dtype, device=self.device)

    def _map_peers(self, shape, dtype):
        """One symmetric allocation: peers[r] is rank r's table, mapped into this rank's address space."""
        import torch.distributed._symmetric_memory as symm
        grp = self.group if self.group is not None else self.dist.group.WORLD
        self.full = symm.empty(shape, dtype=dtype, device=self.device)
        self.full.zero_()
        self._handle = symm.rendezvous(self.full, grp)
        self.peers = [self.full if r == self.rank else self._handle.get_buffer(r, shape, dtype) for r in range(self.world)]

    @property
    def table(self):
        return self.full[: self.n_objects] if self.world * self.per != self.n_objects else self.full

    def local_rows(self, a, b):
        """View of this rank's rows [a, b) (shard-relative): where the kernels write."""
        return self.full[self.start + a: self.start + b]

    def push(self, a, b, after=None):
        """Queue rows [a, b) of this rank's shard (already queued for computation on the current stream) for
        delivery to every other rank.  Returns at once; ``finish`` completes all deliveries."""
        torch, dist = self.torch, self.dist
        if self.world == 1 or b <= a:
            return
        r0, r1 = self.start + a, self.start + b
        if self.cuda:
            ev = after
            if ev is None:
                ev = torch.cuda.Event()
                ev.record()
            self.side.wait_event(ev)
        if self.transport == "p2p":
            with torch.cuda.stream(self.side):
                src = self.full[r0:r1]
                for k in range(1, self.world):             # start with the next rank: spreads the load over the links
                    peer = self.peers[(self.rank + k) % self.world]
                    peer[r0:r1].copy_(src, non_blocking=True)
            return
        # collective: slab [a, b) of EVERY rank lands in a staging block, then moves to its rows.  Ranks whose shard
        # is shorter (the last one) pad with their own rows' storage: the extra rows are never read.
        rows = b - a
        buf = self._n_push & 1
        self._n_push += 1
        if (buf, rows) not in self._stage:                 # two blocks per slab size, used in turn
            self._stage[(buf, rows)] = self.full.new_empty((self.world, rows, self.width))
        stage = self._stage[(buf, rows)]
        src = self.full[self.rank * self.per + a: self.rank * self.per + b]
        if self.cuda:
            with torch.cuda.stream(self.side):
                dist.all_gather_into_tensor(stage.reshape(-1, self.width), src, group=self.group)
                self.full.view(self.world, self.per, self.width)[:, a:b].copy_(stage)
        else:
            dist.all_gather_into_tensor(stage.reshape(-1, self.width), src.contiguous(), group=self.group)
            self.full.view(self.world, self.per, self.width)[:, a:b].copy_(stage)

    def finish(self):
        """Block until every rank's rows have arrived in every table."""
        if self.world == 1:
            if self.cuda:
                self.torch.cuda.current_stream(self.device).synchronize()
            return self.table
        if self.cuda:
            self.side.synchronize()                        # my deliveries are done ...
            self.torch.cuda.current_stream(self.device).synchronize()
        self.dist.barrier(group=self.group)                # ... and so are everybody else's
        return self.table


def slab_bounds(n_rows, slab_rows, tail_rows=None):
    """[(a, b)] covering [0, n_rows) in slabs of slab_rows.  With ``tail_rows`` the last slab is halved again and
    again down to about that size: the delivery of the LAST slab is the only one no computation hides, so it
    should be small (a 16,384-row slab of 720 columns is 94 MB per peer, a 2,048-row one 12 MB)."""
    if n_rows <= 0:
        return []
    slab_rows = max(1, int(slab_rows))
    bounds = [(a, min(n_rows, a + slab_rows)) for a in range(0, n_rows, slab_rows)]
    if tail_rows:
        tail_rows = max(1, int(tail_rows))
        a, b = bounds.pop()
        while b - a > 2 * tail_rows:
            mid = a + (b - a + 1) // 2
            bounds.append((a, mid))
            a = mid
        bounds.append((a, b))
    return bounds


def extract_sharded(extractor, planes, masks=None, sizes=None, hs=None, ws=None, n_objects=None, table=None,
                    slab_objects=16384, group=None, transport=None, finish=True, tail_objects=2048):
    """Extract this rank's shard slab by slab and deliver every slab to all ranks while the next one is computed.

    planes / masks / sizes : this rank's shard, as ``FeatureExtractor.extract_planar`` takes them
                             ([n_local, C, stride] or [n_local, C, Hs, Ws] on the GPU)
    n_objects              : total over all ranks (rank r holds objects shard_range(n_objects, world, r))
    table                  : a ``ShardedTable`` to reuse (its mappings are set up once); made here when None
    Returns the ShardedTable; ``table.table`` is the float64 [n_objects, F] table, complete on every rank once
    ``finish()`` has returned (called here unless finish=False).  All ranks must pass the same slab_objects and
    tail_objects (the last slab is split down to about tail_objects rows so that little is left to deliver
    when the kernels are done; None: no split).
    """
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    assert n_objects is not None
    start, stop, per = shard_range(n_objects, world, rank)
    n_local = stop - start
    assert int(planes.shape[0]) == n_local, "planes must hold exactly this rank's shard"
    C = int(planes.shape[1])
    if table is None:
        table = ShardedTable(n_objects, extractor.row_width(C), device=planes.device, group=group, transport=transport)
    # every rank walks the same slab grid (the collective transport needs matching calls); a rank whose shard is
    # shorter computes fewer rows of its last slabs
    for a, b in slab_bounds(per, slab_objects, tail_objects if world > 1 else None):
        bb = min(b, n_local)
        if bb > a:
            extractor.extract_planar(planes[a:bb], None if masks is None else masks[a:bb],
                                     None if sizes is None else sizes[a:bb], hs=hs, ws=ws,
                                     out=table.local_rows(a, bb))
        if table.transport == "p2p":
            table.push(a, bb)
        else:
            table.push(a, b)
    if finish:
        table.finish()
    return table
