"""Multi-GPU host logic on CPU: partitioners, and the one exchange of the path (lattice broadcast) over a
world-size-2 gloo group."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as tmp

from image_stitcher_b200 import geometry as geo
from image_stitcher_b200 import shard


def test_split_contiguous_is_a_balanced_partition():
    for n in (0, 1, 7, 96, 135, 1152):
        for world in (1, 2, 3, 4, 8):
            parts = [shard.split_contiguous(n, world, r) for r in range(world)]
            assert [i for p in parts for i in p] == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    with pytest.raises(ValueError):
        shard.split_contiguous(4, 2, 2)


def test_wells_round_robin_cover_every_well_once():
    for world in (1, 2, 4, 8):
        owned = sorted(w for r in range(world) for w in shard.wells_for_rank(96, world, r))
        assert owned == list(range(96))
    assert shard.wells_for_rank(384, 8, 3)[:3] == [3, 11, 19] and len(shard.wells_for_rank(384, 8, 3)) == 48


def test_mosaic_pairs_each_pair_owned_once():
    allp = geo.grid_pairs(20, 20)                        # BASELINE.json configs[4]: 760 adjacent pairs
    for world in (2, 8):
        got = [p for r in range(world) for p in shard.mosaic_pairs_for_rank(20, 20, world, r)]
        assert sorted(got) == sorted(allp) and len(got) == 760


def test_fusion_units_cover_the_canvas():
    # configs[4]: 54300-row canvas, 5 planes, 2048-row chunks -> 27 chunk rows x 5 = 135 units
    for world in (1, 2, 8):
        units = [u for r in range(world) for u in shard.fusion_units_for_rank(5, 54300, 2048, world, r)]
        assert len(units) == 135
        for p in range(5):
            rows = sorted((y0, y1) for pp, y0, y1 in units if pp == p)
            assert rows[0][0] == 0 and rows[-1][1] == 54300
            assert all(a[1] == b[0] for a, b in zip(rows, rows[1:]))
    sizes = [len(shard.fusion_units_for_rank(5, 54300, 2048, 8, r)) for r in range(8)]
    assert max(sizes) - min(sizes) <= 1


def test_tiles_for_band_selects_and_rebases():
    tiles = [("a", 0, 0, 0, 0, 0, 10, 0, 0), ("b", 0, 90, 0, 0, 10, 0, 0, 0), ("c", 0, 300, 0, 0, 0, 0, 0, 0)]
    band = shard.tiles_for_band(tiles, 100, 95, 200)
    # a covers rows [0, 90) -> out; b covers [100, 190) -> in, starts 5 rows above the band after re-basing
    assert band == [("b", 0, -5, 0, 0, 10, 0, 0, 0)]
    band = shard.tiles_for_band(tiles, 100, 120, 310)
    assert band == [("b", 0, -30, 0, 0, 30, 0, 0, 0), ("c", 0, 180, 0, 0, 0, 0, 0, 0)]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # rank 0 "solved" the shifts (reference model: once, on the first region); the others start empty
        mine = geo.Lattice((2, -205), (-204, 1), (3, -206), 1, True) if rank == 0 else geo.Lattice()
        got = shard.broadcast_lattice(mine)
        # every rank then places its own wells with the same lattice
        wells = shard.wells_for_rank(6, world, rank)
        place = geo.place_tile(10.0, 20.0, 2048, 2048, [10.0, 11.0], [20.0, 21.0], 0.5, got)
        t = torch.tensor([len(wells)], dtype=torch.int64)
        dist.all_reduce(t)
        out.put((rank, got, wells, place, int(t.item())))
    finally:
        dist.destroy_process_group()


def test_lattice_broadcast_world2_gloo():
    ctx = tmp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted((out.get(timeout=120) for _ in range(2)), key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    expect = geo.Lattice((2, -205), (-204, 1), (3, -206), 1, True)
    assert res[0][1] == expect and res[1][1] == expect
    assert res[0][2] == [0, 2, 4] and res[1][2] == [1, 3, 5]
    assert res[0][3] == res[1][3]
    assert res[0][4] == 6
