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
                 register: bool = True, lattice: Optional[geo.Lattice] = None, partial_upload=False):
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
        # partial_upload (paste mode): a pixel that a later tile overwrites (stitcher_process.py:817) is never read by the
        # fusion kernels, so it need not cross PCIe.  Per tile, the bounding box of what can reach the canvas
        # (geometry.visible_boxes); tiles of the registration channel go up whole (normalize_image scans the whole tile,
        # :844-855).  On the 96-well 3 x 3 plate this takes 9.6 % off the upload -- and measured SLOWER on B200 (r2 call 43:
        # 635 ms per plate against 590 ms with one contiguous copy per well: 27 strided copies of 3.6 KB rows per well do not
        # reach the bus rate of one 302 MB copy), hence off by default.
        self.uploads = None                                  # [(element offset of the tile, x0, y0, x1, y1)] or None = one copy
        if blend == "paste" and partial_upload:
            order = well_fuse_tiles(spec, lambda r, c, ch, z: (r, c, ch, z), lattice)
            boxes = geo.visible_boxes(order, H, W, Hc, Wc)
            ups = []
            for (key, *_), box in zip(order, boxes):
                r, c, ch, z = key
                off = int(np.dot(strides, (r, c, ch, z))) * H * W
                if register and ch == spec.reg_channel and z == 0:
                    box = (0, 0, W, H)
                if box is not None and partial_upload == "rows":     # whole rows only: contiguous copies
                    box = (0, box[1], W, box[3])
                if box is not None:
                    ups.append((off, *box))
            self.uploads = sorted(ups)
            self.upload_bytes = sum((x1 - x0) * (y1 - y0) * 2 for _, x0, y0, x1, y1 in ups)
        else:
            self.upload_bytes = self.well_bytes

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
        if self.uploads is None:
            self.ctx.memcpy_async(lane, self.staging[lane], host_tiles, self.well_bytes, 0)
        else:
            W, H = self.spec.tile_w, self.spec.tile_h
            src0, dst0 = host_tiles.ctypes.data, self.staging[lane]
            for off, x0, y0, x1, y1 in self.uploads:
                o = (off + y0 * W + x0) * 2
                if x0 == 0 and x1 == W:                          # whole rows: one contiguous copy
                    self.ctx.memcpy_async(lane, dst0 + o, src0 + o, (y1 - y0) * W * 2, 0)
                else:
                    self.ctx.memcpy2d_async(lane, dst0 + o, W * 2, src0 + o, W * 2, (x1 - x0) * 2, y1 - y0, 0)
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


class RegionPipeline:
    """Host-buffer fast path of ``StitcherProcess.run`` for uint16 paste jobs that end in OME-Zarr.

    One region per lane, everything enqueued without a host wait: the decoded tiles go up from PINNED staging, the
    region is fused once into a row-major device canvas (from which ``sb_pyramid`` makes the multiscale levels while
    it is still resident) and once in ZARR-CHUNK ORDER straight into a pinned host buffer -- level 0 arrives in the
    layout ``write_ome_zarr_chunked`` dumps with one ``tofile`` per chunk, no re-tiling pass on the host
    (reference: save_region_ome_zarr, stitcher_process.py:1039-1124).  While lane i's copies
    and kernels run, the caller decodes region i+1 and writes region i-1.
    """

    def __init__(self, ctx: _ffi.Context, tile_shape, chunk_hw=(2048, 2048)):
        self.ctx = ctx
        self.H, self.W = int(tile_shape[0]), int(tile_shape[1])
        self.chunk = (int(chunk_hw[0]), int(chunk_hw[1]))
        self.depth = ctx.num_lanes
        self.next = 0
        self.lanes = [dict(tiles_dev=0, tiles_cap=0, canvas_dev=0, canvas_cap=0, pinned_tiles=None, pinned_l0=None,
                           pinned_levels=None, busy=None) for _ in range(self.depth)]

    @staticmethod
    def eligible(dtype, blend, output_format, chunks) -> bool:
        ch, cw = int(chunks[-2]), int(chunks[-1])
        return (np.dtype(dtype) == np.uint16 and blend == _ffi.SB_BLEND_PASTE and str(output_format).endswith(".zarr") and
                cw >= 64 and (cw & (cw - 1)) == 0 and ch % 64 == 0)

    def _grow(self, lane, key, nbytes, pinned=False, dtype=np.uint16):
        st = self.lanes[lane]
        if pinned:
            buf = st[key]
            if buf is None or buf.nbytes < nbytes:
                st[key] = self.ctx.pinned_empty((nbytes // np.dtype(dtype).itemsize,), dtype)
            return st[key]
        if st[key + "_cap"] < nbytes:
            if st[key + "_dev"]:
                self.ctx.sync(lane)
                self.ctx.device_free(st[key + "_dev"])
            st[key + "_dev"] = self.ctx.device_alloc(nbytes)
            st[key + "_cap"] = nbytes
        return st[key + "_dev"]

    def submit(self, job, canvas_shape, n_levels, apply_flatfield, field_c0=0):
        """``job``: ``(plane ndarray, x, y, c, z, crop_t, crop_b, crop_l, crop_r)`` in paste order.  Returns a ticket;
        ``finish(ticket)`` waits for the lane and yields ``(chunked level 0, [levels 1..])`` as arrays over pinned memory
        that stay valid until the lane is used again."""
        lane = self.next
        self.next = (self.next + 1) % self.depth
        self.ctx.sync(lane)
        C, Z, Hc, Wc = (int(v) for v in canvas_shape)
        n = len(job)
        tile_bytes = self.H * self.W * 2
        pinned = self._grow(lane, "pinned_tiles", max(n, 1) * tile_bytes, pinned=True).view(np.uint16)
        dev_tiles = self._grow(lane, "tiles", max(n, 1) * tile_bytes)
        tiles = pinned[:n * self.H * self.W].reshape(n, self.H, self.W)
        dev_job = []
        for i, (plane, x, y, c, z, ct, cb, cl, cr) in enumerate(job):
            tiles[i] = plane                                        # pageable decode buffer -> pinned staging
            dev_job.append((dev_tiles + i * tile_bytes, x, y, c, z, ct, cb, cl, cr))
        if n:
            self.ctx.memcpy_async(lane, dev_tiles, pinned, n * tile_bytes, 0)
        pitch = _ffi.canvas_pitch(Wc)
        dev_canvas = self._grow(lane, "canvas", C * Z * Hc * pitch * 2)
        ch, cw = self.chunk
        ncy, ncx = -(-Hc // ch), -(-Wc // cw)
        l0 = self._grow(lane, "pinned_l0", C * Z * ncy * ncx * ch * cw * 2, pinned=True).view(np.uint16)
        l0 = l0[:C * Z * ncy * ncx * ch * cw].reshape(C * Z, ncy, ncx, ch, cw)
        common = dict(tile_mem=_ffi.SB_MEM_DEVICE, apply_flatfield=apply_flatfield, lane=lane, dtype=_ffi.SB_U16,
                      field_c0=field_c0)           # (a band of one plane of a shared region names its channel here)
        levels = []
        if n_levels > 1:
            # row-major device canvas -> multiscale levels while it is resident (Scaler.nearest, :1061-1062)
            self.ctx.fuse_region(dev_job, (self.H, self.W), (C, Z, Hc, Wc), out=dev_canvas, out_mem=_ffi.SB_MEM_DEVICE, **common)
            shapes = self.ctx.pyramid_shapes((1, C, Z, Hc, Wc), n_levels)
            total = sum(p * a * b for p, a, b in shapes)
            lv = self._grow(lane, "pinned_levels", total * 2, pinned=True).view(np.uint16)[:total]
            self.ctx.pyramid((1, C, Z, Hc, Wc), n_levels, src=None, dtype=_ffi.SB_U16, out=lv, lane=lane)
            off = 0
            for p, a, b in shapes:
                levels.append(lv[off:off + p * a * b].reshape(1, C, Z, a, b))
                off += p * a * b
        # level 0 in zarr-chunk order, straight to the host
        self.ctx.fuse_region(dev_job, (self.H, self.W), (C, Z, Hc, Wc), out=l0, out_mem=_ffi.SB_MEM_HOST,
                             layout=_ffi.SB_LAYOUT_CHUNKED, chunk=(ch, cw), **common)
        return dict(lane=lane, l0=l0, levels=levels, shape=(C, Z, Hc, Wc))

    def finish(self, ticket):
        self.ctx.sync(ticket["lane"])
        return ticket["l0"], ticket["levels"]

    def close(self):
        self.ctx.sync(-1)
        for st in self.lanes:
            for key in ("tiles", "canvas"):
                if st[key + "_dev"]:
                    self.ctx.device_free(st[key + "_dev"])
                    st[key + "_dev"], st[key + "_cap"] = 0, 0
