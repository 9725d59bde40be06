"""CPU tests (gloo, world_size 2) of the N>1 host logic: object sharding and the all-gather of the
per-rank feature blocks.  The kernels are not involved; each rank fabricates the rows of its shard
from the object index, so the gathered table can be checked exactly."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import imfeat_b200 as imf
from imfeat_b200 import distributed as D


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _rows(start, stop, width):
    idx = torch.arange(start, stop, dtype=torch.float64)[:, None]
    return idx * 1000.0 + torch.arange(width, dtype=torch.float64)[None, :]


def _worker(rank, world, port, n_objects, width, chunk, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        start, stop, per = D.shard_range(n_objects, world, rank)
        local = _rows(start, stop, width)
        full = D.gather_table(local, n_objects, chunk_rows=chunk)
        ok = bool(torch.equal(full, _rows(0, n_objects, width)))
        q.put((rank, ok, tuple(full.shape)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_objects,chunk", [(7, None), (8, None), (1, None), (13, 3), (64, 16)])
def test_gather_table_gloo_world2(n_objects, chunk):
    world, width = 2, 23
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_objects, width, chunk, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok, shape in res:
        assert ok, "rank %d gathered a wrong table" % rank
        assert shape == (n_objects, width)


class _FakeExtractor:
    """Stands in for FeatureExtractor on CPU tensors: row i of the table is a function of the object index
    carried in planes[i, 0, 0]."""

    def __init__(self, width):
        self.width = width

    def row_width(self, c):
        return self.width

    def extract_planar(self, planes, masks=None, sizes=None, hs=None, ws=None, out=None):
        idx = planes[:, 0, 0].to(torch.float64)[:, None]
        out.copy_(idx * 1000.0 + torch.arange(self.width, dtype=torch.float64)[None, :])
        return out


def _sharded_worker(rank, world, port, n_objects, width, slab, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        start, stop, per = D.shard_range(n_objects, world, rank)
        planes = torch.arange(start, stop, dtype=torch.float64).reshape(-1, 1, 1)
        tab = D.extract_sharded(_FakeExtractor(width), planes, n_objects=n_objects, slab_objects=slab)
        ok = tab.transport == "collective" and bool(torch.equal(tab.table, _rows(0, n_objects, width)))
        # the table object is reusable: a second extraction into the same buffers
        tab2 = D.extract_sharded(_FakeExtractor(width), planes, n_objects=n_objects, slab_objects=slab, table=tab)
        ok = ok and tab2 is tab and bool(torch.equal(tab.table, _rows(0, n_objects, width)))
        q.put((rank, ok, tuple(tab.table.shape)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_objects,slab", [(7, 2), (8, 4), (1, 16), (13, 3), (64, 64), (65, 16)])
def test_extract_sharded_slab_pipeline_gloo_world2(n_objects, slab):
    """The product's slab pipeline (kernels of slab k+1 while slab k is delivered) with the collective
    transport on CPU tensors: every rank ends up with the full table, also with ragged last shards."""
    world, width = 2, 23
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_sharded_worker, args=(r, world, port, n_objects, width, slab, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok, shape in res:
        assert ok, "rank %d assembled a wrong table" % rank
        assert shape == (n_objects, width)


def test_slab_bounds():
    assert D.slab_bounds(0, 4) == []
    assert D.slab_bounds(10, 4) == [(0, 4), (4, 8), (8, 10)]
    assert D.slab_bounds(4, 100) == [(0, 4)]
    # tapered tail: the last slab is halved down to about tail_rows; the cover stays exact
    assert D.slab_bounds(16, 16, tail_rows=2) == [(0, 8), (8, 12), (12, 16)]
    assert D.slab_bounds(10, 4, tail_rows=1) == [(0, 4), (4, 8), (8, 10)]
    b = D.slab_bounds(125000, 16384, tail_rows=2048)
    assert b[0] == (0, 16384) and b[-1][1] == 125000 and all(x[1] == y[0] for x, y in zip(b, b[1:]))
    assert b[-1][1] - b[-1][0] <= 4096 and max(y - x for x, y in b) == 16384


def test_shard_ranges_cover_everything():
    for n in (0, 1, 7, 8, 1000003):
        for world in (1, 2, 4, 8):
            spans = [D.shard_range(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0]
            assert all(stop - start <= per for start, stop, per in spans)


def test_balanced_ranges_by_pixel_count():
    rng = np.random.default_rng(0)
    w = (rng.integers(16, 129, 500) * rng.integers(16, 129, 500)).tolist()
    for world in (2, 4, 8):
        spans = D.balanced_ranges(w, world)
        assert spans[0][0] == 0 and spans[-1][1] == len(w)
        loads = [sum(w[a:b]) for a, b in spans]
        assert max(loads) < 1.15 * sum(w) / world
        for a, b in zip(spans, spans[1:]):
            assert a[1] == b[0]


def test_ablation_index_tables():
    from imfeat_b200 import ablation
    lists = ablation.loco_channel_lists(12)
    assert lists.shape == (12, 11) and all(k not in lists[k] for k in range(12))
    src = ablation.permutation_sources(100, 12, seed=42)
    assert src.shape == (12, 100, 12)
    perm0 = np.random.default_rng(42).permutation(100)
    assert (src[0][:, 0] == perm0).all()
    for k in range(12):
        others = np.delete(src[k], k, axis=1)
        assert (others == np.arange(100)[:, None]).all()
        assert sorted(src[k][:, k].tolist()) == list(range(100))
    groups = ablation.group_importances_by_channel(imf.feature_columns(12), np.full(276, 0.02), 0.01)
    assert sorted(groups) == list(range(1, 13)) and all(len(v) == 23 for v in groups.values())
