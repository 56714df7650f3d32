"""Host-buffer pipeline: tiles of one region in (pinned) host memory -> registered shifts + fused
host canvas, overlapped across the context's lanes (H2D of region i+1, kernels of region i and D2H
of region i-1 run concurrently on different streams / copy engines).

This is the public call the end-to-end benchmark times; it is also what the reference-facing
``StitcherProcess.stitch_region`` uses when it is asked for several regions.
"""
from __future__ import annotations

from typing import List, Optional

import numpy as np

from . import _ffi
from . import geometry as geo
from .plate import FusePlan, PlateSpec, well_fuse_tiles, well_pairs


class WellPipeline:
    def __init__(self, ctx: _ffi.Context, spec: PlateSpec, *, apply_flatfield: bool, blend: str = "paste",
                 register: bool = True, lattice: Optional[geo.Lattice] = None):
        self.ctx, self.spec = ctx, spec
        self.depth = ctx.num_lanes
        self.register = register
        self.next = 0
        H, W = spec.tile_h, spec.tile_w
        self.tile_bytes = H * W * 2
        self.well_bytes = spec.tiles_per_well * self.tile_bytes
        Wc, Hc = spec.canvas_size() if lattice is None else geo.canvas_size(
            W, H, *spec.stage_positions(), spec.pixel_size_um, lattice)
        self.canvas_shape = (spec.channels, spec.num_z, Hc, Wc)
        self.ov = spec.strip_overlaps()
        self.staging, self.plans, self.pairs = [], [], []
        self.inflight = [None] * self.depth
        self.last_lane = None
        shape = (spec.rows, spec.cols, spec.channels, spec.num_z)
        strides = np.array([spec.cols * spec.channels * spec.num_z, spec.channels * spec.num_z, spec.num_z, 1])
        for lane in range(self.depth):
            base = ctx.device_alloc(self.well_bytes)
            self.staging.append(base)

            def ptr(r, c, ch, z, base=base):
                return base + int(np.dot(strides, (r, c, ch, z))) * self.tile_bytes

            # the host canvas pointer is patched into the job at submit time
            self.plans.append(FusePlan(ctx, well_fuse_tiles(spec, ptr, lattice), (H, W), self.canvas_shape, 0,
                                       tile_mem=_ffi.SB_MEM_DEVICE, out_mem=_ffi.SB_MEM_HOST,
                                       apply_flatfield=apply_flatfield, blend=_ffi.BLEND_MODES[blend],
                                       blend_ov=self.ov))
            self.pairs.append(well_pairs(spec, ptr)[0])

    def submit(self, host_tiles: np.ndarray, host_out: np.ndarray):
        """``host_tiles``: uint16 [rows, cols, C, Z, H, W] (C-contiguous, ideally pinned).
        ``host_out``: uint16 (1, C, Z, Hc, Wc).  Everything is enqueued on the next lane -- upload, registration
        (``sb_register_pairs_async``), fusion, download -- and the call returns immediately, so the upload of the next
        region overlaps this one's kernels and the previous one's download.  Returns a ``PendingRegistration`` whose
        ``get()`` yields the pair results once the lane has been synchronised (next reuse of the lane, or ``drain``)."""
        assert host_tiles.dtype == np.uint16 and host_tiles.flags.c_contiguous
        assert host_out.shape[-4:] == self.canvas_shape and host_out.flags.c_contiguous
        lane = self.next
        self.next = (self.next + 1) % self.depth
        self.ctx.sync(lane)                                  # the lane's previous region is complete (results landed)
        if self.inflight[lane] is not None:
            self.inflight[lane].mark_synced()
        # uploads of consecutive regions run back to back (this lane's upload waits for the previous lane's upload,
        # not for its kernels or download), so the two bus directions stay busy at the same time
        if self.last_lane is not None:
            self.ctx.lane_wait_mark(lane, self.last_lane)
        self.ctx.memcpy_async(lane, self.staging[lane], host_tiles, self.well_bytes, 0)
        self.ctx.lane_mark(lane)
        self.last_lane = lane
        pending = None
        if self.register:
            pending = self.ctx.register_pairs_async(self.pairs[lane], (self.spec.tile_h, self.spec.tile_w), self.ov[0],
                                                    self.ov[1], mem=_ffi.SB_MEM_DEVICE, lane=lane)
        self.inflight[lane] = pending
        plan = self.plans[lane]
        plan.job.out = host_out.ctypes.data
        plan.run(lane)
        return pending

    def drain(self):
        self.ctx.sync(-1)
        for p in self.inflight:
            if p is not None:
                p.mark_synced()

    def close(self):
        self.drain()
        for p in self.staging:
            self.ctx.device_free(p)
        self.staging = []
