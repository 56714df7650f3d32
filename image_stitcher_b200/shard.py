"""Multi-GPU partitioning of the hot path (one process per GPU).

The path shards into independent units -- there is NO data-path collective:

* registration: by tile pair, whole wells per rank so that every tile's min/max scan happens once;
  inside one large mosaic, by grid-row bands of the pair list;
* fusion: by well / region; inside one large mosaic by ``(plane, chunk-row)`` units, each of which only
  needs the tiles that intersect its rows.

The only datum shared between ranks is the solved lattice of the reference's registration model
(``h_shift``, ``v_shift``, ``h_shift_rev``, ``h_shift_rev_odd`` -- 7 ints, computed once on the first region,
stitcher_process.py:1975-1976): ``broadcast_lattice`` sends it from rank 0 with ``torch.distributed``
(NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

from . import geometry as geo


def split_contiguous(n_units: int, world: int, rank: int) -> range:
    """Balanced contiguous block of ``range(n_units)`` for ``rank`` (sizes differ by at most one)."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, extra = divmod(n_units, world)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def wells_for_rank(n_wells: int, world: int, rank: int) -> List[int]:
    """Round-robin wells ``rank, rank + world, ...``: neighbouring wells (similar content / file locality)
    spread over the GPUs, every well's pairs and canvas on exactly one rank."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError(f"bad rank/world {rank}/{world}")
    return list(range(rank, n_wells, world))


def mosaic_pairs_for_rank(n_rows: int, n_cols: int, world: int, rank: int):
    """Adjacent pairs of ONE big rows x cols mosaic, split by grid-row bands.  A pair belongs to the band of its
    reference tile's row, so every pair is owned once; a boundary row's tiles are scanned by two ranks."""
    rows = split_contiguous(n_rows, world, rank)
    return [p for p in geo.grid_pairs(n_rows, n_cols) if p[1][0] in rows]


def fusion_units_for_rank(n_planes: int, canvas_h: int, chunk_h: int, world: int, rank: int) -> List[Tuple[int, int, int]]:
    """``(plane, y0, y1)`` output bands of one big canvas: the ``(plane, chunk-row)`` product split evenly."""
    n_cy = -(-canvas_h // chunk_h)
    units = [(p, cy) for p in range(n_planes) for cy in range(n_cy)]
    out = []
    for i in split_contiguous(len(units), world, rank):
        p, cy = units[i]
        out.append((p, cy * chunk_h, min((cy + 1) * chunk_h, canvas_h)))
    return out


def tiles_for_band(tiles: Sequence[tuple], tile_h: int, y0: int, y1: int) -> List[tuple]:
    """Tiles ``(px, x, y, c, z, crop_t, crop_b, crop_l, crop_r)`` whose cropped rows intersect ``[y0, y1)``, re-based
    so that the band starts at canvas row 0 (paste order preserved).  Rows above the band are removed by raising
    ``crop_t`` (the library rejects negative canvas positions); rows below it fall off the band canvas.  Exact for
    paste mode (a crop only hides pixels); blend modes weight by distance to the cropped edge, so shard those by
    region instead."""
    out = []
    for t in tiles:
        px, x, y, c, z, ct, cb, cl, cr = t
        if y + ct < y1 and y + tile_h - cb > y0:
            out.append((px, x, y - y0, c, z, max(ct, y0 - y), cb, cl, cr))
    return out


def broadcast_lattice(lattice: geo.Lattice, src: int = 0, device=None) -> geo.Lattice:
    """Send the solved shifts from ``src`` to every rank (the one exchange of the path; 7 int64 values)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return lattice
    vals = [*lattice.h_shift, *lattice.v_shift, *lattice.h_shift_rev, int(lattice.h_shift_rev_odd), int(lattice.s_pattern)]
    t = torch.tensor(vals if dist.get_rank() == src else [0] * 8, dtype=torch.int64, device=device)
    dist.broadcast(t, src=src)
    v = [int(x) for x in t.tolist()]
    return geo.Lattice((v[0], v[1]), (v[2], v[3]), (v[4], v[5]), v[6], bool(v[7]))
