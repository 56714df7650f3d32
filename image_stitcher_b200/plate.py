"""Synthetic tiled plates resident on the device + the batched registration / fusion driver.

Used by bench.py, ``__graft_entry__.smoke()`` and the full-size GPU tests.  torch is plumbing
here (device memory, streams, the random generator); every timed operation is a call into
libstitchb200 through ``_ffi``.

Plate recipe: per well a "world" image (power-law noise, gaussian
blurred, scaled into uint16) for the registration channel; tiles are cut on the lattice
``step = round(0.9 * W)`` with an integer per-well stage drift (the ground truth the
registration must recover) and get independent sensor noise.  The other channels carry
cheap pseudo-random content (fusion cost does not depend on pixel values).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np

from . import _ffi
from . import geometry as geo


@dataclass
class PlateSpec:
    wells: int = 96
    rows: int = 3
    cols: int = 3
    tile_h: int = 2048
    tile_w: int = 2048
    channels: int = 4
    num_z: int = 1
    overlap: float = 0.10
    jitter: int = 3
    reg_channel: int = 1          # "Fluorescence 488 nm Ex" is the 2nd of 405/488/561/638 in sorted order
    pixel_size_um: float = 0.5
    pixel_binning: int = 2
    seed: int = 0

    @property
    def step(self) -> Tuple[int, int]:
        return int(round((1 - self.overlap) * self.tile_h)), int(round((1 - self.overlap) * self.tile_w))

    @property
    def tiles_per_well(self) -> int:
        return self.rows * self.cols * self.channels * self.num_z

    def stage_positions(self):
        """Nominal stage coordinates in mm, as coordinates.csv would report them."""
        sy, sx = self.step
        xs = [10.0 + c * sx * self.pixel_size_um / 1000.0 for c in range(self.cols)]
        ys = [20.0 + r * sy * self.pixel_size_um / 1000.0 for r in range(self.rows)]
        return xs, ys

    def canvas_size(self) -> Tuple[int, int]:
        xs, ys = self.stage_positions()
        return geo.canvas_size(self.tile_w, self.tile_h, xs, ys, self.pixel_size_um, None)

    def strip_overlaps(self) -> Tuple[int, int]:
        xs, ys = self.stage_positions()
        return geo.strip_overlaps(self.tile_w, self.tile_h, xs, ys, self.pixel_size_um, self.pixel_binning)


@dataclass
class Plate:
    spec: PlateSpec
    pool: "object"                    # torch int16 tensor [wells, rows, cols, C, Z, H, W] (uint16 bit patterns)
    truth: List[Dict]                 # per well: expected (dy, dx) of horizontal / vertical pairs
    flat: Optional["object"] = None   # torch float32 [C, H, W]


def vignette_torch(h, w, strength, shift, device):
    import torch
    yy = torch.linspace(0, 1, h, device=device, dtype=torch.float64)[:, None] - 0.5 - shift[0]
    xx = torch.linspace(0, 1, w, device=device, dtype=torch.float64)[None, :] - 0.5 - shift[1]
    ff = 1.0 - strength * (yy * yy + xx * xx) / 0.5
    return (ff / ff.mean()).to(torch.float32)


def make_plate(spec: PlateSpec, device="cuda", with_flat=True, well_ids=None) -> Plate:
    """Generate the plate directly in device memory (never on the timed path).

    Every well is seeded by ``(spec.seed, well id)``, so a rank that owns wells ``well_ids`` of a larger plate (strong
    scaling: ``shard.wells_for_rank``) generates exactly the wells a single GPU would hold under those ids."""
    import torch
    g = torch.Generator(device=device)
    well_ids = list(range(spec.wells)) if well_ids is None else [int(w) for w in well_ids]
    assert len(well_ids) == spec.wells, "spec.wells is the number of wells THIS plate holds"
    H, W = spec.tile_h, spec.tile_w
    sy, sx = spec.step
    j = spec.jitter
    pad = 2 * j * max(spec.rows, spec.cols) + 8
    wh = H + (spec.rows - 1) * sy + 2 * pad
    ww = W + (spec.cols - 1) * sx + 2 * pad
    pool = torch.empty((spec.wells, spec.rows, spec.cols, spec.channels, spec.num_z, H, W), dtype=torch.int16,
                       device=device)
    flat = None
    if with_flat:
        flat = torch.stack([vignette_torch(H, W, 0.35, (0.04 * (c + 1), -0.03 * (c + 1)), device)
                            for c in range(spec.channels)])
    k1 = torch.tensor([0.054, 0.244, 0.403, 0.244, 0.054], device=device)      # gaussian, sigma = 1
    truth = []
    for wl, wid in enumerate(well_ids):
        g.manual_seed(spec.seed * 1000003 + wid)
        rng = np.random.default_rng([spec.seed, wid])
        col_jx, row_jy, col_jy, row_jx = (int(v) for v in rng.integers(-j, j + 1, 4)) if j else (0, 0, 0, 0)
        truth.append({"h": (col_jy, -(W - sx) + col_jx), "v": (-(H - sy) + row_jy, row_jx)})
        u = torch.rand((1, 1, wh, ww), generator=g, device=device) ** 6
        u = torch.nn.functional.conv2d(u, k1.view(1, 1, 1, 5), padding=(0, 2))
        u = torch.nn.functional.conv2d(u, k1.view(1, 1, 5, 1), padding=(2, 0))[0, 0]
        world = 400.0 + 40000.0 * u / u.max()
        for r in range(spec.rows):
            for c in range(spec.cols):
                wx = pad + c * (sx + col_jx) + r * row_jx + j * max(spec.rows, spec.cols)
                wy = pad + r * (sy + row_jy) + c * col_jy + j * max(spec.rows, spec.cols)
                base = world[wy:wy + H, wx:wx + W]
                for ch in range(spec.channels):
                    for z in range(spec.num_z):
                        if ch == spec.reg_channel:
                            img = base * (1.0 - 0.08 * z) + 30.0 * torch.randn((H, W), generator=g, device=device)
                        else:
                            # other channels: the same structure, different gain (keeps the flat-field
                            # division busy on realistic values) -- not used for registration
                            img = base * (0.3 + 0.15 * ch) + 50.0
                        if flat is not None:
                            img = img * flat[ch]
                        pool[wl, r, c, ch, z] = img.clamp_(0, 65535).to(torch.int32).to(torch.int16)
    return Plate(spec, pool, truth, flat)


class FusePlan:
    """A prebuilt ``sb_fuse_job`` (ctypes) for one region, reusable every step without Python marshalling."""

    def __init__(self, ctx: _ffi.Context, tiles, tile_shape, canvas_shape, out_ptr, *, tile_mem, out_mem,
                 apply_flatfield, blend=_ffi.SB_BLEND_PASTE, blend_ov=(0, 0), layout=_ffi.SB_LAYOUT_ROWMAJOR,
                 out_row_pitch=0, chunk=(0, 0)):
        self.ctx = ctx
        n = len(tiles)
        self.arr = (_ffi.SbTile * max(n, 1))()
        for i, (px, x, y, c, z, ct, cb, cl, cr) in enumerate(tiles):
            self.arr[i] = _ffi.SbTile(_ffi._ptr(px), int(x), int(y), int(c), int(z), int(ct), int(cb), int(cl), int(cr))
        self.job = _ffi.SbFuseJob(self.arr, n, int(tile_shape[0]), int(tile_shape[1]), _ffi.SB_U16, tile_mem,
                                  int(canvas_shape[0]), int(canvas_shape[1]), int(canvas_shape[2]),
                                  int(canvas_shape[3]), int(bool(apply_flatfield)), int(blend), int(blend_ov[0]),
                                  int(blend_ov[1]), _ffi._ptr(out_ptr), out_mem, layout, int(out_row_pitch),
                                  int(chunk[0]), int(chunk[1]))

    def run(self, lane: int = 0):
        rc = self.ctx.lib.sb_fuse_region(self.ctx.handle, C.byref(self.job), lane)
        if rc != 0:
            self.ctx._check(rc, "sb_fuse_region")


class FuseBatchPlan:
    """The regions of a plate in ONE ``sb_fuse_regions`` call (prebuilt ctypes job array; the per-region plans are kept
    alive because the jobs point into their tile arrays)."""

    def __init__(self, ctx: _ffi.Context, plans: List[FusePlan]):
        self.ctx, self.plans = ctx, list(plans)
        self.jobs = (_ffi.SbFuseJob * max(len(self.plans), 1))(*[p.job for p in self.plans])

    def run(self, lane: int = 0):
        rc = self.ctx.lib.sb_fuse_regions(self.ctx.handle, self.jobs, len(self.plans), lane)
        if rc != 0:
            self.ctx._check(rc, "sb_fuse_regions")


class RegisterPlan:
    """A prebuilt ``sb_register_job`` (ctypes) for a fixed pair list, reusable every step: ``run()`` is the bare
    ``sb_register_pairs`` call (building 1152 ``sb_pair`` structs and as many result dicts in Python costs milliseconds
    per call -- more than some of the kernels); ``results()`` converts the last run's records on demand."""

    def __init__(self, ctx: _ffi.Context, pairs, tile_shape, max_overlap_x, max_overlap_y, *, mem, upsample_factor=10,
                 precision=_ffi.SB_PREC_AUTO, lane=0, dtype=None):
        self.ctx, self.lane = ctx, int(lane)
        self.arr, self.res, self.job = ctx._register_job(list(pairs), tile_shape, max_overlap_x, max_overlap_y, mem,
                                                         upsample_factor, precision, lane, dtype)

    def run(self):
        rc = self.ctx.lib.sb_register_pairs(self.ctx.handle, C.byref(self.job), self.res)
        if rc != 0:
            self.ctx._check(rc, "sb_register_pairs")
        return self

    def run_async(self):
        """``sb_register_pairs_async``: the records are valid after ``ctx.sync(lane)``."""
        rc = self.ctx.lib.sb_register_pairs_async(self.ctx.handle, C.byref(self.job), self.res)
        if rc != 0:
            self.ctx._check(rc, "sb_register_pairs_async")
        return self

    def results(self):
        return _ffi.Context._pair_dicts(self.res)


def well_fuse_tiles(spec: PlateSpec, tile_ptr, lattice: Optional[geo.Lattice] = None):
    """sb_tile tuples of one well in the reference's paste order.

    ``tile_ptr(r, c, ch, z)`` -> address (device) or array (host).  The reference pastes in sorted
    file-name order ``<region>_<fov>_<z>_<channel>`` (stitcher_process.py:283-288): lexicographic in
    the fov *string*, then z, then channel name (channel index order == sorted name order here).
    """
    xs, ys = spec.stage_positions()
    fovs = sorted(range(spec.rows * spec.cols), key=lambda f: str(f))
    out = []
    for fov in fovs:
        r, c = divmod(fov, spec.cols)
        p = geo.place_tile(xs[c], ys[r], spec.tile_w, spec.tile_h, xs, ys, spec.pixel_size_um, lattice)
        for z in range(spec.num_z):
            for ch in range(spec.channels):
                out.append((tile_ptr(r, c, ch, z), p.x, p.y, ch, z, p.crop_t, p.crop_b, p.crop_l, p.crop_r))
    return out


def well_pairs(spec: PlateSpec, tile_ptr, z: int = 0):
    """All adjacent pairs of one well on the registration channel: ``(ref, mov, dir)`` + kinds."""
    pairs, kinds = [], []
    for kind, (r0, c0), (r1, c1) in geo.grid_pairs(spec.rows, spec.cols):
        d = _ffi.SB_DIR_HORIZONTAL if kind == "h" else _ffi.SB_DIR_VERTICAL
        pairs.append((tile_ptr(r0, c0, spec.reg_channel, z), tile_ptr(r1, c1, spec.reg_channel, z), d))
        kinds.append(kind)
    return pairs, kinds
