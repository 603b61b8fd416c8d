"""Host-side logic of the multi-GPU sharding (eggfusion_b200/parallel.py) on CPU: tile partition, surfel ranges,
and the one collective (reduce-scatter of the screen-gradient block) with world_size 2 over gloo."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from eggfusion_b200 import parallel as par


def test_tile_partition_is_a_partition():
    for world in (1, 2, 3, 4, 8):
        ty, tx = 68, 120
        masks = [par.tile_partition(ty, tx, world, r) for r in range(world)]
        total = torch.stack(masks).sum(0)
        assert torch.equal(total, torch.ones_like(total))
        assert all(m.dtype == torch.int32 for m in masks)


@pytest.mark.parametrize("layout", ["bands", "rows"])
def test_cost_balanced_partition(layout):
    ty, tx, world = 40, 30, 4
    rng = np.random.default_rng(0)
    costs = torch.from_numpy(rng.pareto(1.5, size=(ty, tx)) * 100)
    masks = [par.tile_partition(ty, tx, world, r, costs, layout) for r in range(world)]
    assert torch.equal(torch.stack(masks).sum(0), torch.ones((ty, tx), dtype=torch.int32))
    loads = torch.tensor([float((costs * m).sum()) for m in masks])
    naive = torch.tensor([float((costs * par.tile_partition(ty, tx, world, r)).sum()) for r in range(world)])
    assert loads.max() <= naive.max() + 1e-9
    assert loads.max() / loads.mean() < (1.05 if layout == "bands" else 1.25)
    if layout == "bands":
        # every rank's tiles are ONE contiguous run of the row-major sequence, in rank order
        owner = torch.stack(masks).argmax(0).reshape(-1)
        assert bool((owner[1:] >= owner[:-1]).all())
    with pytest.raises(ValueError):
        par.tile_partition(ty, tx, world, 0, costs, "diagonal")


def test_band_partition_degenerate_costs():
    """All-zero costs (nothing rendered yet) still give every rank a run; one tile holding all the cost goes to one rank."""
    ty, tx, world = 9, 7, 8
    masks = [par.tile_partition(ty, tx, world, r, torch.zeros(ty * tx)) for r in range(world)]
    assert torch.equal(torch.stack(masks).sum(0), torch.ones((ty, tx), dtype=torch.int32))
    assert all(int(m.sum()) >= ty * tx // world - 1 for m in masks)
    spike = torch.zeros(ty * tx)
    spike[20] = 1e9
    masks = [par.tile_partition(ty, tx, world, r, spike) for r in range(world)]
    assert torch.equal(torch.stack(masks).sum(0), torch.ones((ty, tx), dtype=torch.int32))


def test_surfel_ranges_cover_exactly():
    for P in (0, 1, 7, 1000, 1_000_003):
        for world in (1, 2, 4, 8):
            rows = par.padded_rows(P, world)
            assert rows % world == 0 and rows >= P and rows - P < world * par.CHUNK_ALIGN
            assert world == 1 or (rows // world) % par.CHUNK_ALIGN == 0
            spans = [par.surfel_range(P, world, r) for r in range(world)]
            covered = 0
            for first, count in spans:
                assert first == min(P, covered) or count == 0
                covered += count
            assert covered == P


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, P, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rows = par.padded_rows(P, world)
        g = torch.Generator().manual_seed(100 + rank)
        block = torch.zeros((rows, 16))
        block[:P] = torch.randn((P, 16), generator=g)
        mine = par.reduce_scatter_rows(block, None)
        first, count = par.surfel_range(P, world, rank)
        out[rank] = (first, count, mine.clone())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("P", [10, 1001])
def test_reduce_scatter_rows_gloo_world2(P):
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, P, out), nprocs=world, join=True)
    rows = par.padded_rows(P, world)
    full = torch.zeros((rows, 16))
    for r in range(world):
        g = torch.Generator().manual_seed(100 + r)
        full[:P] += torch.randn((P, 16), generator=g)
    chunk = rows // world
    covered = 0
    for r in range(world):
        first, count, mine = out[r]
        assert mine.shape == (chunk, 16)
        assert torch.allclose(mine, full[r * chunk:(r + 1) * chunk], atol=1e-6)
        assert first == r * chunk or count == 0
        covered += count
    assert covered == P


def _gather_worker(rank, world, port, P, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        chunk = par.padded_rows(P, world) // world
        first, count = par.surfel_range(P, world, rank)
        # every rank starts from the same replica and "updates" only its owned rows
        t = torch.arange(P * 3, dtype=torch.float32).view(P, 3).clone()
        t[first:first + count] += 1000.0 * (rank + 1)
        par.all_gather_rows(t, first, count, chunk, world)
        out[rank] = t.clone()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("P", [7, 1000])
def test_all_gather_rows_gloo_world2(P):
    """The second half of SURVEY 8e: after the owners updated their rows, one all-gather makes every replica whole."""
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_gather_worker, args=(world, port, P, out), nprocs=world, join=True)
    want = torch.arange(P * 3, dtype=torch.float32).view(P, 3).clone()
    for r in range(world):
        first, count = par.surfel_range(P, world, r)
        want[first:first + count] += 1000.0 * (r + 1)
    for r in range(world):
        assert torch.equal(out[r], want)


def test_sharded_api_surface():
    """The autograd-level sharded rasterizer mirrors GaussianRasterizer's argument checks; the exchange falls back to
    NCCL / gloo collectives when peer memory is unavailable (here: no process group at all -> world 1)."""
    sh = par.ShardedSplat()
    assert sh.world == 1 and sh.rank == 0
    assert par.make_exchange(100, "cpu") is None
    r = par.ShardedRasterizer(raster_settings=None, sharder=sh)
    with pytest.raises(Exception, match="excatly one"):
        r(means3D=torch.zeros(1, 3), opacities=torch.zeros(1, 1))
    assert r.owned_range(10) == (0, 10)
