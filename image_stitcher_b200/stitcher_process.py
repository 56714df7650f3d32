"""``StitcherProcess`` -- host-side mirror of the reference's orchestrator for the hot path.

Same class name, constructor, queue protocol and method names as the reference's
``stitcher_process.StitcherProcess`` (stitcher_process.py:61-2037), so code written against the
reference (its CLI/GUI, ``save_region_test.py``-style harnesses that inject state by attribute
assignment) drives this class unchanged.  The L1 methods -- ``normalize_image``,
``calculate_horizontal_shift`` / ``calculate_vertical_shift`` / ``calculate_shifts``,
``apply_flatfield_correction``, ``stitch_region`` -- call libstitchb200 through ``_ffi``; there is
no NumPy fallback (no CUDA device -> RuntimeError).  Everything else here is host glue written
from the reference's *behaviour* (file layout, ordering and rounding rules are cited inline).

Out of scope: BaSiC flat-field *fitting*, OME-TIFF / bioio / aicsimageio /
pyvips writers, pyramid merges, HCS plate merges.  ``get_flatfields`` uses BaSiCPy when it is
installed and otherwise the GPU's robust estimator (``sb_estimate_flatfield``, an extension with its own
definition -- or provide the fields with ``set_flatfields``); ``.ome.zarr`` output goes through the minimal
NGFF writer in ``ome_zarr_writer``.
"""
from __future__ import annotations

import json
import math
import os
import sys
import time
from multiprocessing import Process
from queue import Empty
from typing import Dict, List, Optional

import numpy as np

from . import _ffi
from . import geometry as geo
from .stitcher_parameters import StitchingParameters

IMAGE_SUFFIXES = (".bmp", ".tiff", "tif", "jpg", "jpeg", "png")     # as written in the reference (:285)
CHANNEL_COLORS = (("405", 0x0000FF), ("488", 0x00FF00), ("561", 0xFFCF00), ("638", 0xFF0000), ("730", 0x770000),
                  ("_B", 0x0000FF), ("_G", 0x00FF00), ("_R", 0xFF0000))


def read_image(path: str) -> np.ndarray:
    """Decode one tile (what ``dask_imread(path)[0]`` yields in the reference, :338/:731/:913)."""
    import cv2
    img = cv2.imread(path, cv2.IMREAD_UNCHANGED)
    if img is None:
        raise FileNotFoundError(path)
    if img.ndim == 3 and img.shape[2] == 3:
        img = img[:, :, ::-1]          # OpenCV decodes BGR; the reference's readers give RGB
    return np.ascontiguousarray(img)


class StitcherProcess(Process):
    def __init__(self, params: StitchingParameters, progress_queue, status_queue, complete_queue, stop_event):
        super().__init__()
        self.progress_queue, self.status_queue = progress_queue, status_queue
        self.complete_queue, self.stop_event = complete_queue, stop_event
        self.params = params
        self.input_folder = params.input_folder
        self.output_folder = params.stitched_folder
        self.output_format = params.output_format
        self.merge_timepoints = getattr(params, "merge_timepoints", False)
        self.merge_hcs_regions = getattr(params, "merge_hcs_regions", False)
        self.per_timepoint_region_output_template = os.path.join(
            self.output_folder, "{timepoint}_stitched", "{region}_stitched" + self.output_format)
        self.apply_flatfield = params.apply_flatfield
        self.use_registration = params.use_registration
        self.registration_channel = params.registration_channel if self.use_registration else ""
        self.registration_z_level = params.registration_z_level if self.use_registration else 0
        self.dynamic_registration = getattr(params, "dynamic_registration", False)
        self.scan_pattern = params.scan_pattern
        self.blend_mode = getattr(params, "blend_mode", "paste")
        self.upsample_factor = int(getattr(params, "upsample_factor", 10))
        self.registration_precision = getattr(params, "registration_precision", "auto")
        self.placement = getattr(params, "placement", "lattice")
        self.visualize_registration = bool(getattr(params, "visualize_registration", False))
        self.tile_positions: Dict[tuple, Dict[tuple, tuple]] = {}     # (t, region) -> {(x_mm, y_mm): (x_px, y_px)}
        self.device = int(getattr(params, "device", 0))
        self.decode_threads = max(1, min(16, os.cpu_count() or 1))
        # multi-GPU: worker `rank` of `world` stitches regions rank, rank + world, ... on its own device.  Every worker
        # registers the same first region itself (2-3 pairs), so all of them hold the same lattice without any exchange.
        self.rank, self.world = int(getattr(params, "rank", 0)), max(1, int(getattr(params, "world", 1)))
        # One region on several GPUs: with fewer regions than workers (or split_regions=True) every worker takes part in
        # EVERY region and fuses its share of the (plane, chunk-row) bands of the canvas (shard.fusion_units_for_rank).
        self.split_regions = bool(getattr(params, "split_regions", False))
        self._ctx: Optional[_ffi.Context] = None
        self._pyramids = {}                               # (timepoint, region) -> (n_levels, GPU-made levels 1..)
        self._flat_dirty = True
        self.init_stitching_parameters()

    # ------------------------------------------------------------------ state / IPC (reference :146-230)
    def init_stitching_parameters(self):
        self.pixel_size_um = None
        self.acquisition_params = None
        self.timepoints: List[str] = []
        self.regions: List[str] = []
        self.channel_names: List[str] = []
        self.monochrome_channels: List[str] = []
        self.monochrome_colors: List[int] = []
        self.num_z = self.num_c = self.num_t = 1
        self.input_height = self.input_width = 0
        self.num_pyramid_levels = 5
        self.flatfields: Dict[int, np.ndarray] = {}
        self.acquisition_metadata: Dict[tuple, dict] = {}
        self.dtype = np.uint16
        self.chunks = (1, 1, 1, 2048, 2048)
        self.h_shift = (0, 0)
        if self.scan_pattern == "S-Pattern":
            self.h_shift_rev = (0, 0)
            self.h_shift_rev_odd = 0
        self.v_shift = (0, 0)
        self.x_positions = set()
        self.y_positions = set()
        self.pixel_binning = 1

    def emit_progress(self, current: int, total: int):
        if self.progress_queue is None:
            print(f"PROGRESS: {(current, total)}")
        else:
            self.progress_queue.put(("progress", (current, total)))

    def emit_status(self, status: str, is_saving: bool = False):
        if self.status_queue is None:
            print(f"STATUS: {status}")
        else:
            self.status_queue.put(("status", (status, is_saving)))

    def emit_complete(self, output_path: str, dtype):
        if self.complete_queue is None:
            print("COMPLETE:")
        else:
            self.complete_queue.put(("complete", (output_path, dtype)))

    def check_stop(self):
        """Cooperative stop, polled per region (the reference polls per tile; a region is one kernel launch)."""
        if self.stop_event is not None and self.stop_event.is_set():
            print("Stop event detected, terminating process...")
            self.cleanup()
            sys.exit(0)

    def cleanup(self):
        for q in (self.progress_queue, self.status_queue, self.complete_queue):
            if q is None:
                continue
            try:
                while not q.empty():
                    q.get_nowait()
            except (Empty, OSError):
                pass
        if self._ctx is not None:
            self._ctx.close()
            self._ctx = None
        self.emit_status("Process Stopped...")

    # ------------------------------------------------------------------ CUDA context (lazy: created in the worker)
    @property
    def ctx(self) -> _ffi.Context:
        if self._ctx is None:
            self._ctx = _ffi.Context(self.device)        # raises without a CUDA device: no CPU fallback
            self._flat_dirty = True
        return self._ctx

    def set_flatfields(self, flatfields: Dict[int, np.ndarray]):
        """Install per-channel-index flatfields (what ``get_flatfields`` produces, :524)."""
        self.flatfields = dict(flatfields)
        self._flat_dirty = True

    def _sync_fields(self):
        if self._flat_dirty:
            self.ctx.clear_fields()
            for c, ff in self.flatfields.items():
                self.ctx.set_flatfield(int(c), np.asarray(ff))
            self._flat_dirty = False

    # ------------------------------------------------------------------ metadata (reference :232-421)
    def get_timepoints(self):
        self.timepoints = sorted((d for d in os.listdir(self.input_folder)
                                  if d.isdigit() and os.path.isdir(os.path.join(self.input_folder, d))), key=int)
        return self.timepoints

    def extract_acquisition_parameters(self):
        with open(os.path.join(self.input_folder, "acquisition parameters.json")) as fh:
            self.acquisition_params = json.load(fh)

    def get_pixel_size(self):
        p = self.acquisition_params
        self.pixel_binning = p.get("pixel_binning", 1)
        obj_focal_length_mm = p["objective"]["tube_lens_f_mm"] / p["objective"]["magnification"]
        self.pixel_size_um = p["sensor_pixel_size_um"] / (p["tube_lens_mm"] / obj_focal_length_mm)

    def parse_acquisition_metadata(self):
        """File names ``<region>_<fov>_<z>_<channel>.<ext>`` matched to coordinates.csv rows (:261-371).
        Dict insertion order == sorted file names: it is the paste order of ``stitch_region``."""
        import pandas as pd
        self.acquisition_metadata = {}
        regions, channels = set(), set()
        max_z = max_fov = 0
        for timepoint in self.timepoints:
            folder = os.path.join(self.input_folder, str(timepoint))
            try:
                df = pd.read_csv(os.path.join(folder, "coordinates.csv"))
            except FileNotFoundError:
                print(f"Warning: coordinates.csv not found for timepoint {timepoint}")
                continue
            cols = list(df.columns)
            ix, iy, iz = cols.index("x (mm)"), cols.index("y (mm)"), cols.index("z (um)")
            rows = {(str(r[cols.index("region")]), int(r[cols.index("fov")]), int(r[cols.index("z_level")])): r
                    for r in reversed(list(df.itertuples(index=False, name=None)))}     # first match wins (:303)
            files = sorted(f for f in os.listdir(folder)
                           if f.endswith(IMAGE_SUFFIXES) and not f.startswith(".") and "focus_camera" not in f)
            for name in files:
                parts = name.split("_", 3)
                region, fov, z_level = parts[0], int(parts[1]), int(parts[2])
                channel = os.path.splitext(parts[3])[0].replace("_", " ").replace("full ", "full_")
                row = rows.get((region, fov, z_level))
                if row is None:
                    print(f"Warning: No coordinates for {name}")
                    continue
                self.acquisition_metadata[(int(timepoint), region, fov, z_level, channel)] = {
                    "filepath": os.path.join(folder, name), "x": row[ix], "y": row[iy], "z": row[iz],
                    "channel": channel, "z_level": z_level, "region": region, "fov_idx": fov, "t": int(timepoint)}
                regions.add(region)
                channels.add(channel)
                max_z, max_fov = max(max_z, z_level), max(max_fov, fov)
        self.regions, self.channel_names = sorted(regions), sorted(channels)
        self.num_t, self.num_z, self.num_fovs_per_region = len(self.timepoints), max_z + 1, max_fov + 1
        first_key = next(iter(self.acquisition_metadata))
        first = read_image(self.acquisition_metadata[first_key]["filepath"])
        self.dtype = first.dtype
        self.input_height, self.input_width = first.shape[:2]
        self.monochrome_channels = []
        t0, r0, f0, z0, _ = first_key
        for channel in self.channel_names:
            img = read_image(self.acquisition_metadata[(t0, r0, f0, z0, channel)]["filepath"])
            if img.ndim == 3 and img.shape[2] == 3:
                base = channel.split("_")[0]
                self.monochrome_channels += [f"{base}_R", f"{base}_G", f"{base}_B"]
            else:
                self.monochrome_channels.append(channel)
        self.num_c = len(self.monochrome_channels)
        self.monochrome_colors = [self.get_channel_color(n) for n in self.monochrome_channels]

    def get_region_data(self, t, region):
        t = int(t)
        data = {k: v for k, v in self.acquisition_metadata.items() if k[0] == t and k[1] == region}
        if not data:
            raise ValueError(f"No data found for timepoint {t}, region {region}")
        return data

    def get_channel_color(self, channel_name):
        for key, color in CHANNEL_COLORS:
            if key in channel_name:
                return color
        return 0xFFFFFF

    def get_rows_and_columns(self):
        return sorted({r[0] for r in self.regions}), sorted({r[1:] for r in self.regions})

    # ------------------------------------------------------------------ geometry (reference :423-503)
    def _lattice(self) -> Optional[geo.Lattice]:
        if not self.use_registration:
            return None
        s = self.scan_pattern == "S-Pattern"
        return geo.Lattice(tuple(self.h_shift), tuple(self.v_shift), tuple(getattr(self, "h_shift_rev", (0, 0))),
                           int(getattr(self, "h_shift_rev_odd", 0)), s)

    def calculate_output_dimensions(self, timepoint, region):
        data = self.get_region_data(int(timepoint), region)
        self.x_positions = sorted({v["x"] for v in data.values()})
        self.y_positions = sorted({v["y"] for v in data.values()})
        width, height = geo.canvas_size(self.input_width, self.input_height, self.x_positions, self.y_positions,
                                        self.pixel_size_um, self._lattice())
        max_dim = 1
        if len(self.regions) > 1:
            rows, cols = self.get_rows_and_columns()
            max_dim = max(len(rows), len(cols))
        self.num_pyramid_levels = geo.pyramid_levels(width, height, len(self.regions), max_dim)
        return width, height

    def init_output(self, timepoint, region):
        """The reference returns a lazy zero canvas; here the canvas is produced whole by ``stitch_region``."""
        width, height = self.calculate_output_dimensions(timepoint, region)
        return np.zeros((1, self.num_c, self.num_z, height, width), dtype=self.dtype)

    # ------------------------------------------------------------------ flat-field estimation (out of scope)
    def get_flatfields(self):
        """Per-channel flat-fields from a random sample of tiles (:505-571: at most 32 tiles per timepoint, stop above
        48).  With BaSiCPy installed the fit is the reference's ``BaSiC(get_darkfield=False, smoothness_flatfield=1)``;
        without it (this image) the sample goes to the GPU's robust estimator ``sb_estimate_flatfield`` -- an
        extension with its own definition (oracle/flatfield_ref.py), NOT a reimplementation of BaSiC."""
        import random
        try:
            from basicpy import BaSiC
        except ImportError:
            BaSiC = None
            self.emit_status("BaSiCPy is not installed: flat-fields from the GPU median estimator (not BaSiC)")

        def process_images(images, channel_name):
            if images.size == 0:
                print(f"Warning: No images found for channel {channel_name}")
                return
            if images.ndim == 4:                           # (N, Z, Y, X): every plane is one sample
                images = images.reshape((-1,) + images.shape[-2:])
            if images.ndim != 3:
                raise ValueError("Images must be 3 or 4-dimensional array, with dimension of (T, Y, X) or (T, Z, Y, X). "
                                 f"Got shape {images.shape}")
            channel_index = self.monochrome_channels.index(channel_name)
            if BaSiC is not None:
                basic = BaSiC(get_darkfield=False, smoothness_flatfield=1)
                basic.fit(images)
                self.flatfields[channel_index] = basic.flatfield
            else:
                self.flatfields[channel_index] = self.ctx.estimate_flatfield(list(images[:128]))
            self.emit_progress(channel_index + 1, self.num_c)

        self.emit_progress(0, self.num_c)
        for channel in self.channel_names:
            self.check_stop()
            self.emit_status(f"Calculating Flatfield... ({channel})")
            images = []
            for t in self.timepoints:
                paths = [tile["filepath"] for key, tile in self.acquisition_metadata.items()
                         if tile["channel"] == channel and key[0] == int(t)]
                if not paths:
                    print(f"Warning: No images found for channel {channel} at timepoint {t}")
                    continue
                random.shuffle(paths)
                images.extend(read_image(p) for p in paths[:min(32, len(paths))])
                if len(images) > 48:
                    break
            if not images:
                print(f"Warning: No images found for channel {channel} across all timepoints")
                continue
            images = np.array(images)
            if images.ndim == 4 and images.shape[-1] == 3:  # (N, Y, X, 3) RGB tiles: one field per colour plane
                base = channel.split("_")[0]
                for i, suffix in enumerate(("_R", "_G", "_B")):
                    process_images(np.ascontiguousarray(images[..., i]), base + suffix)
            else:
                process_images(images, channel)
        self._flat_dirty = True

    # ------------------------------------------------------------------ registration (reference :573-737, :844-855)
    def normalize_image(self, img):
        """Whole-tile min/max stretch in float64 with truncating cast (:844-855) -- ``sb_normalize``."""
        return self.ctx.normalize(np.ascontiguousarray(img, dtype=self._pixel_np()))

    def _pixel_np(self):
        """uint8 or uint16 -- the dtype of the first image, like the reference's ``self.dtype`` (:340)."""
        dt = np.dtype(self.dtype)
        if dt not in (np.dtype(np.uint8), np.dtype(np.uint16)):
            raise RuntimeError(f"pixel dtype {dt} is not supported (uint8 and uint16 are)")
        return dt

    def _precision(self) -> int:
        return {"auto": _ffi.SB_PREC_AUTO, "float32": _ffi.SB_PREC_F32, "float64": _ffi.SB_PREC_F64}[self.registration_precision]

    def _register(self, pairs, max_x_overlap, max_y_overlap):
        shape = tuple(pairs[0][0].shape)
        dt = self._pixel_np()
        for a, b, _ in pairs:
            # the C ABI reads tile_h * tile_w pixels from every pointer: a tile of another shape (different ROI, truncated
            # file, RGB plane) must not reach it
            if np.ndim(a) != 2 or tuple(a.shape) != shape or tuple(b.shape) != shape:
                raise ValueError(f"registration tiles must be 2-D and share one shape: {np.shape(a)} / {np.shape(b)} vs {shape}")
        job = [(np.ascontiguousarray(a, dtype=dt), np.ascontiguousarray(b, dtype=dt), d) for a, b, d in pairs]
        if self.visualize_registration:
            self._visualize_pairs(job, max_x_overlap, max_y_overlap)
        return self.ctx.register_pairs(job, shape, max_x_overlap, max_y_overlap, upsample_factor=self.upsample_factor,
                                       precision=self._precision())

    def _visualize_pairs(self, job, max_x_overlap, max_y_overlap):
        """The reference's debug side effect (:681, :704): the normalised overlap strips of the LAST horizontal and the
        last vertical pair, side by side, as <out>/horizontal.png and <out>/vertical.png."""
        last = {}
        for a, b, d in job:
            last[d] = (a, b)
        for d, (a, b) in last.items():
            na, nb = self.normalize_image(a), self.normalize_image(b)
            if d == _ffi.SB_DIR_HORIZONTAL:
                m = int(a.shape[0] * 0.25)
                self.visualize_image(na[m:-m, -max_x_overlap:], nb[m:-m, :max_x_overlap], "horizontal")
            else:
                m = int(a.shape[1] * 0.25)
                self.visualize_image(na[-max_y_overlap:, m:-m], nb[:max_y_overlap, m:-m], "vertical")

    def visualize_image(self, img1, img2, title):
        """(:857-881) the two strips stacked (side by side for 'horizontal') as an 8-bit PNG in the output folder."""
        try:
            import cv2
            img1, img2 = np.asarray(img1), np.asarray(img2)
            combined = np.hstack((img1, img2)) if title == "horizontal" else np.vstack((img1, img2))
            combined8 = (combined / np.iinfo(self.dtype).max * 255).astype(np.uint8)
            os.makedirs(self.output_folder, exist_ok=True)
            cv2.imwrite(f"{self.output_folder}/{title}.png", combined8)
            print(f"Saved {title}.png successfully")
        except Exception as e:
            print(f"Error in visualize_image: {e}")

    def calculate_horizontal_shift(self, img_left, img_right, max_overlap):
        """(:664-685) -> ``(round(shift[0]), round(shift[1] - strip_width))``."""
        r = self._register([(img_left, img_right, _ffi.SB_DIR_HORIZONTAL)], max_overlap, max_overlap)[0]
        return r["dy"], r["dx"]

    def calculate_vertical_shift(self, img_top, img_bot, max_overlap):
        """(:687-708) -> ``(round(shift[0] - strip_height), round(shift[1]))``."""
        r = self._register([(img_top, img_bot, _ffi.SB_DIR_VERTICAL)], max_overlap, max_overlap)[0]
        return r["dy"], r["dx"]

    def get_tile(self, t, region, x, y, channel, z_level):
        for value in self.get_region_data(int(t), str(region)).values():
            if value["x"] == x and value["y"] == y and value["channel"] == channel and value["z_level"] == z_level:
                try:
                    return read_image(value["filepath"])
                except FileNotFoundError:
                    print(f"Warning: Tile file not found: {value['filepath']}")
                    return None
        print(f"Warning: No matching tile found for region {region}, x={x}, y={y}, channel={channel}, z={z_level}")
        return None

    def calculate_shifts(self, t, region):
        """The 2 (3 for S-Pattern) centre pairs of the first region, one batched GPU call (:573-662)."""
        self.h_shift = (0, 0)
        self.v_shift = (0, 0)
        if not self.registration_channel or self.registration_channel not in self.channel_names:
            if self.registration_channel:
                print(f"Warning: Registration channel '{self.registration_channel}' not found")
            self.registration_channel = self.channel_names[0]
        self.emit_status("Calculating Registration Shifts...")
        self.calculate_output_dimensions(int(t), region)
        xs, ys = list(self.x_positions), list(self.y_positions)
        ov_x, ov_y = geo.strip_overlaps(self.input_width, self.input_height, xs, ys, self.pixel_size_um, self.pixel_binning)
        s_pat = self.scan_pattern == "S-Pattern"
        plan, rev_odd = geo.center_pairs(xs, ys, s_pat)
        if self.dynamic_registration:
            # extension behind the reference's declared-but-unused flag (stitcher_parameters.py:24): register EVERY
            # adjacent pair of the region in one batch and keep the per-direction median (lower median, an observed value)
            plan = []
            for kind, (r0, c0), (r1, c1) in geo.grid_pairs(len(ys), len(xs)):
                if kind == "h" and s_pat and r0 % 2 == int(rev_odd):
                    kind = "h_rev"
                plan.append((kind, (xs[c0], ys[r0]), (xs[c1], ys[r1])))
        pairs, kinds = [], []
        for kind, a_xy, b_xy in plan:
            a = self.get_tile(t, region, a_xy[0], a_xy[1], self.registration_channel, self.registration_z_level)
            b = self.get_tile(t, region, b_xy[0], b_xy[1], self.registration_channel, self.registration_z_level)
            if a is None or b is None:
                print(f"Warning: Missing tiles for {kind} shift calculation in region {region}.")
                continue
            pairs.append((a, b, _ffi.SB_DIR_VERTICAL if kind == "v" else _ffi.SB_DIR_HORIZONTAL))
            kinds.append(kind)
        if pairs:
            found = {"h": [], "v": [], "h_rev": []}
            for kind, r in zip(kinds, self._register(pairs, ov_x, ov_y)):
                found[kind].append((r["dy"], r["dx"]))
            self.registration_results = found

            def pick(lst):
                dys, dxs = sorted(p[0] for p in lst), sorted(p[1] for p in lst)
                return dys[(len(dys) - 1) // 2], dxs[(len(dxs) - 1) // 2]
            if found["h"]:
                self.h_shift = pick(found["h"])
            if found["v"]:
                self.v_shift = pick(found["v"])
            if found["h_rev"]:
                self.h_shift_rev = pick(found["h_rev"])
                self.h_shift_rev_odd = rev_odd
        print(f"Calculated Shifts - Horizontal: {self.h_shift}, Vertical: {self.v_shift}")

    def register_region_global(self, t, region):
        """Extension (``placement='global'``): register EVERY adjacent pair of this region in one GPU batch and solve
        the tile origins by least squares (``geometry.solve_positions``).  The reference has neither per-pair
        registration nor a global solve; its lattice model stays the default."""
        if not self.registration_channel or self.registration_channel not in self.channel_names:
            self.registration_channel = self.channel_names[0]
        self.calculate_output_dimensions(int(t), region)
        xs, ys = list(self.x_positions), list(self.y_positions)
        ov_x, ov_y = geo.strip_overlaps(self.input_width, self.input_height, xs, ys, self.pixel_size_um, self.pixel_binning)
        index = {(x, y): r * len(xs) + c for r, y in enumerate(ys) for c, x in enumerate(xs)}
        tiles = {}
        for (x, y) in index:
            tiles[(x, y)] = self.get_tile(t, region, x, y, self.registration_channel, self.registration_z_level)
        plan, job = [], []
        for kind, (r0, c0), (r1, c1) in geo.grid_pairs(len(ys), len(xs)):
            a, b = tiles[(xs[c0], ys[r0])], tiles[(xs[c1], ys[r1])]
            if a is None or b is None:
                continue
            plan.append((kind, index[(xs[c0], ys[r0])], index[(xs[c1], ys[r1])]))
            job.append((a, b, _ffi.SB_DIR_VERTICAL if kind == "v" else _ffi.SB_DIR_HORIZONTAL))
        res = self._register(job, ov_x, ov_y) if job else []
        pos = geo.solve_positions(len(index), self.input_width, self.input_height, plan, [(r["dy"], r["dx"]) for r in res],
                                  weights=[max(float(r["peak"]), 1e-3) for r in res])
        self.tile_positions[(int(t), region)] = {xy: pos[i] for xy, i in index.items()}
        return self.tile_positions[(int(t), region)]

    # ------------------------------------------------------------------ fusion (reference :739-956)
    def apply_flatfield_correction(self, tile, channel_idx):
        """``(tile / flatfield).clip(0, max).astype(dtype)`` (:828-842) -- ``sb_flatfield_apply``."""
        if channel_idx not in self.flatfields:
            return tile
        self._sync_fields()
        return self.ctx.flatfield_apply(int(channel_idx), np.ascontiguousarray(tile, dtype=self._pixel_np()))

    def _tile_planes(self, tile: np.ndarray, channel: str):
        """(:739-769): mono -> one plane; H x W x 3 -> <ch>_R/_G/_B planes; 1 x H x W -> squeezed."""
        if tile.ndim == 2:
            return [(self.monochrome_channels.index(channel), tile)]
        if tile.ndim == 3 and tile.shape[2] == 3:
            base = channel.split("_")[0]
            return [(self.monochrome_channels.index(f"{base}_{c}"), np.ascontiguousarray(tile[:, :, i]))
                    for i, c in enumerate("RGB")]
        if tile.ndim == 3 and tile.shape[0] == 1:
            return [(self.monochrome_channels.index(channel), tile[0])]
        raise ValueError(f"Unexpected tile shape: {tile.shape}")

    def _decode_region(self, timepoint, region):
        """Decode every tile of a region on a thread pool (cv2 releases the GIL).  Returns ``[(key, info, tile | None,
        exc)]`` in dict order -- the sorted-file-name paste order of the reference (:283-288, :908)."""
        from concurrent.futures import ThreadPoolExecutor
        data = self.get_region_data(int(timepoint), region)

        def _load(item):
            key, info = item
            try:
                return key, info, read_image(info["filepath"]), None
            except Exception as exc:                       # the reference reports and skips (:912-916)
                return key, info, None, exc
        with ThreadPoolExecutor(max_workers=self.decode_threads) as pool:
            return list(pool.map(_load, data.items()))

    def _region_job(self, timepoint, region, loaded=None):
        """Geometry of one region: the paste-ordered ``sb_tile`` tuples and the canvas shape (:883-942)."""
        data = self.get_region_data(int(timepoint), region)
        width, height = self.calculate_output_dimensions(timepoint, region)
        lattice = self._lattice()
        xs, ys = list(self.x_positions), list(self.y_positions)
        solved = None
        if self.use_registration and self.placement == "global":
            solved = self.tile_positions.get((int(timepoint), region)) or self.register_region_global(timepoint, region)
            width = max(p[0] for p in solved.values()) + self.input_width
            height = max(p[1] for p in solved.values()) + self.input_height
        self.emit_status(f"Stitching... (Timepoint:{timepoint} Region:{region})")
        self.check_stop()
        job = []
        if loaded is None:
            loaded = self._decode_region(timepoint, region)
        for key, info, tile, exc in loaded:
            if tile is None:
                self.emit_status(f"Error Loading Image {info['filepath']}: {exc}")
                continue
            if solved is not None:
                p = geo.Placement(*solved[(info["x"], info["y"])])
            else:
                p = geo.place_tile(info["x"], info["y"], self.input_width, self.input_height, xs, ys,
                                   self.pixel_size_um, lattice)
            for c, plane in self._tile_planes(tile, key[4]):
                if plane.shape != (self.input_height, self.input_width):
                    # the C ABI reads input_height x input_width pixels per tile: skip what does not have them
                    self.emit_status(f"Error Loading Image {info['filepath']}: shape {plane.shape} != "
                                     f"{(self.input_height, self.input_width)}")
                    continue
                plane = np.ascontiguousarray(plane, dtype=self._pixel_np())
                job.append((plane, p.x, p.y, c, key[3], p.crop_t, p.crop_b, p.crop_l, p.crop_r))
        return job, (self.num_c, self.num_z, height, width), len(data)

    def stitch_region(self, timepoint, region, loaded=None):
        """One region -> ``(1, C, Z, Hc, Wc)`` NumPy canvas, ONE ``sb_fuse_region`` call (:883-956).  ``loaded`` takes the
        result of an earlier ``_decode_region`` (``run`` decodes the next region while this one is fused and saved)."""
        start = time.time()
        try:
            job, (_, _, height, width), n_data = self._region_job(timepoint, region, loaded)
            xs, ys = list(self.x_positions), list(self.y_positions)
            out = np.empty((1, self.num_c, self.num_z, height, width), dtype=self._pixel_np())
            if self.apply_flatfield:
                self._sync_fields()
            ov = geo.strip_overlaps(self.input_width, self.input_height, xs, ys, self.pixel_size_um, 2) \
                if len(xs) > 1 and len(ys) > 1 else (0, 0)
            blend = _ffi.BLEND_MODES[self.blend_mode]
            if blend != _ffi.SB_BLEND_PASTE:               # blending needs the overlaps: no seam crops
                job = [t[:5] + (0, 0, 0, 0) for t in job]
            self.ctx.fuse_region(job, (self.input_height, self.input_width), (self.num_c, self.num_z, height, width),
                                 out=out, apply_flatfield=self.apply_flatfield, blend=blend, blend_ov=ov)
            if self.num_pyramid_levels > 1 and str(self.output_format).endswith(".zarr"):
                # multiscale levels from the canvas while it is still on the device (Scaler.nearest, :1061-1062)
                self._pyramids[(int(timepoint), region)] = (self.num_pyramid_levels, self.ctx.pyramid(
                    out.shape, self.num_pyramid_levels, dtype=_ffi._pixel_dtype(out)))
                while len(self._pyramids) > 4:             # callers that never save: do not hoard levels (dicts keep insertion order)
                    self._pyramids.pop(next(iter(self._pyramids)))
            self.emit_progress(n_data, n_data)
            print(f"(Timepoint:{timepoint}, Region:{region}) Complete Stitching in {time.time() - start:.1f}s\n")
            return out
        except Exception as exc:
            if self.status_queue is not None:
                self.status_queue.put(("error", f"Error stitching region {region}: {exc}"))
            raise

    def place_single_channel_tile(self, stitched_region, tile, x_pixel, y_pixel, z_level, channel_idx, t):
        """Per-tile form of the paste (:771-826), kept for callers that place tiles one by one.  The flat-field
        division runs on the GPU; the slice assignment itself is a host copy into the caller's array."""
        if stitched_region.ndim != 5:
            raise ValueError(f"Unexpected stitched_region shape: {stitched_region.shape}. Expected 5D array (t, c, z, y, x).")
        if self.apply_flatfield:
            tile = self.apply_flatfield_correction(tile, channel_idx)
        if self.use_registration:
            # row_index / col_index are set by the caller, as in the reference's stitch_region (:920-921)
            xs, ys = list(self.x_positions), list(self.y_positions)
            p = geo.place_tile(xs[self.col_index], ys[self.row_index], self.input_width, self.input_height, xs, ys,
                               self.pixel_size_um, self._lattice())
            tile = tile[p.crop_t:tile.shape[0] - p.crop_b, p.crop_l:tile.shape[1] - p.crop_r]
            x_pixel += p.crop_l
            y_pixel += p.crop_t
        y_end = min(y_pixel + tile.shape[0], stitched_region.shape[3])
        x_end = min(x_pixel + tile.shape[1], stitched_region.shape[4])
        stitched_region[0, channel_idx, z_level, y_pixel:y_end, x_pixel:x_end] = tile[:y_end - y_pixel, :x_end - x_pixel]

    def place_tile(self, stitched_region, tile, x_pixel, y_pixel, z_level, channel, t):
        for c, plane in self._tile_planes(tile, channel):
            self.place_single_channel_tile(stitched_region, plane, x_pixel, y_pixel, z_level, c, 0)

    # ------------------------------------------------------------------ output
    # ------------------------------------------------------------------ one region on several GPUs
    def _channel_planes(self, channel: str):
        """Canvas channel indices a file of ``channel`` feeds (one for mono, three for RGB) -- without decoding it."""
        if channel in self.monochrome_channels:
            return [self.monochrome_channels.index(channel)]
        base = channel.split("_")[0]
        return [self.monochrome_channels.index(f"{base}_{c}") for c in "RGB" if f"{base}_{c}" in self.monochrome_channels]

    def band_groups(self, height: int):
        """This worker's share of a region's canvas: ``(plane, y0, y1)`` with consecutive chunk rows of a plane merged
        (``shard.fusion_units_for_rank`` splits the (plane, chunk-row) product evenly over the workers)."""
        from .shard import fusion_units_for_rank
        groups = []
        for p, y0, y1 in fusion_units_for_rank(self.num_c * self.num_z, height, int(self.chunks[-2]), self.world, self.rank):
            if groups and groups[-1][0] == p and groups[-1][2] == y0:
                groups[-1] = (p, groups[-1][1], y1)
            else:
                groups.append((p, y0, y1))
        return groups

    def _band_job(self, timepoint, region, plane, y0, y1):
        """Paste-ordered job of the tiles of plane ``(c, z)`` whose kept rows reach canvas rows ``[y0, y1)``, re-based to
        the band (``shard.tiles_for_band``).  Only those files are decoded."""
        from concurrent.futures import ThreadPoolExecutor
        from .shard import tiles_for_band
        c, z = divmod(int(plane), self.num_z)
        data = self.get_region_data(int(timepoint), region)
        lattice = self._lattice()
        xs, ys = list(self.x_positions), list(self.y_positions)
        want = []
        for key, info in data.items():                     # dict order == paste order
            if key[3] != z or c not in self._channel_planes(key[4]):
                continue
            p = geo.place_tile(info["x"], info["y"], self.input_width, self.input_height, xs, ys, self.pixel_size_um, lattice)
            if p.y + p.crop_t < y1 and p.y + self.input_height - p.crop_b > y0:
                want.append((key, info, p))

        def _load(item):
            key, info, p = item
            try:
                return item, read_image(info["filepath"]), None
            except Exception as exc:
                return item, None, exc
        with ThreadPoolExecutor(max_workers=self.decode_threads) as pool:
            loaded = list(pool.map(_load, want))
        job = []
        for (key, info, p), tile, exc in loaded:
            if tile is None:
                self.emit_status(f"Error Loading Image {info['filepath']}: {exc}")
                continue
            for ci, arr in self._tile_planes(tile, key[4]):
                if ci != c:
                    continue
                if arr.shape != (self.input_height, self.input_width):
                    self.emit_status(f"Error Loading Image {info['filepath']}: shape {arr.shape} != "
                                     f"{(self.input_height, self.input_width)}")
                    continue
                job.append((np.ascontiguousarray(arr, dtype=self._pixel_np()), p.x, p.y, 0, 0,
                            p.crop_t, p.crop_b, p.crop_l, p.crop_r))
        return tiles_for_band(job, self.input_height, y0, y1), c, z

    def _save_band(self, timepoint, region, pipe, ticket, c, z, y0, full_shape, n_levels):
        from .ome_zarr_writer import write_ome_zarr_band
        l0, levels = pipe.finish(ticket)
        path = self.per_timepoint_region_output_template.format(timepoint=timepoint, region=region)
        dz = self.acquisition_params.get("dz(um)", 1.0) if self.acquisition_params else 1.0
        return write_ome_zarr_band(path, l0, levels, plane=(c, z), row0=y0, full_shape=full_shape, chunk_hw=self.chunks[-2:],
                                   n_levels=n_levels, pixel_size_um=self.pixel_size_um, dz_um=dz,
                                   channel_names=self.monochrome_channels, channel_colors=self.monochrome_colors,
                                   name=f"{region}_t{timepoint}")

    def _run_bands(self, pipe, writer):
        """Every worker walks every (timepoint, region) and fuses ITS bands: decode of the band's tiles, one
        ``RegionPipeline`` submission per (plane, row range) -- level 0 in chunk order + the band's pyramid levels --
        and a writer thread that drops the chunks into the shared OME-Zarr.  No exchange between the workers: chunk
        files of level 0 have one owner; rows of the coarser levels go through memory maps of shared chunk files."""
        last_path, lane_writes = "", {}
        for timepoint in self.timepoints:
            for region in self.regions:
                self.check_stop()
                width, height = self.calculate_output_dimensions(timepoint, region)
                full_shape = (self.num_c, self.num_z, height, width)
                n_levels = self.num_pyramid_levels
                if self.apply_flatfield:
                    self._sync_fields()
                groups = self.band_groups(height)
                self.emit_status(f"Stitching... (Timepoint:{timepoint} Region:{region}) bands {len(groups)} of worker "
                                 f"{self.rank}/{self.world}")
                for gi, (plane, y0, y1) in enumerate(groups):
                    self.check_stop()
                    job, c, z = self._band_job(timepoint, region, plane, y0, y1)
                    busy = lane_writes.pop(pipe.next, None)
                    if busy is not None:
                        last_path = busy.result()          # the lane's pinned buffers are free again
                    ticket = pipe.submit(job, (1, 1, y1 - y0, width), n_levels, self.apply_flatfield, field_c0=c)
                    self.emit_progress(gi + 1, len(groups))
                    lane_writes[ticket["lane"]] = writer.submit(self._save_band, timepoint, region, pipe, ticket, c, z, y0,
                                                                full_shape, n_levels)
        for fut in lane_writes.values():
            last_path = fut.result()
        return last_path

    def save_region_ome_zarr(self, timepoint, region, stitched_region, num_levels=None):
        from .ome_zarr_writer import write_ome_zarr
        path = self.per_timepoint_region_output_template.format(timepoint=timepoint, region=region)
        dz = self.acquisition_params.get("dz(um)", 1.0) if self.acquisition_params else 1.0
        stitched_region = np.asarray(stitched_region)
        n_levels = self.num_pyramid_levels if num_levels is None else num_levels
        made_for, levels = self._pyramids.pop((int(timepoint), region), (0, None))
        if levels is not None and (made_for < n_levels or levels[0].shape[-2:] !=
                                   ((stitched_region.shape[-2] + 1) // 2, (stitched_region.shape[-1] + 1) // 2)):
            levels = None                                  # not the canvas stitch_region produced: slice on the host
        write_ome_zarr(path, stitched_region, pixel_size_um=self.pixel_size_um, dz_um=dz,
                       channel_names=self.monochrome_channels, channel_colors=self.monochrome_colors,
                       num_levels=n_levels, chunks=self.chunks, levels=levels, name=f"{region}_t{timepoint}")
        return path

    def _save_ticket(self, timepoint, region, pipe, ticket):
        """Writer-thread half of the fast path: wait for the lane, dump level 0 chunk by chunk, levels 1.. as made on the GPU."""
        from .ome_zarr_writer import write_ome_zarr_chunked
        l0, levels = pipe.finish(ticket)
        path = self.per_timepoint_region_output_template.format(timepoint=timepoint, region=region)
        dz = self.acquisition_params.get("dz(um)", 1.0) if self.acquisition_params else 1.0
        write_ome_zarr_chunked(path, l0, ticket["shape"], self.chunks[-2:], pixel_size_um=self.pixel_size_um, dz_um=dz,
                               channel_names=self.monochrome_channels, channel_colors=self.monochrome_colors,
                               name=f"{region}_t{timepoint}", levels=levels)
        return path

    def run(self):
        """Same sequence as the reference's ``run`` (:1959-2037)."""
        stime = time.time()
        try:
            # requests this build does not serve are refused BEFORE any decode / GPU work
            if not self.output_format.endswith(".zarr"):
                raise RuntimeError("OME-TIFF output relies on the reference's third-party writers "
                                   "(out of scope, SURVEY.md section 2); use .ome.zarr")
            if self.merge_timepoints or self.merge_hcs_regions:
                self.emit_status("Warning: merge_timepoints / merge_hcs_regions are not implemented (the reference marks its "
                                 "merges 'not ready', SURVEY.md section 2.3): writing one OME-Zarr per region instead")
            self.emit_status("Extracting Acquisition Metadata...")
            self.get_timepoints()
            self.extract_acquisition_parameters()
            self.get_pixel_size()
            self.parse_acquisition_metadata()
            os.makedirs(self.output_folder, exist_ok=True)
            last_path = ""
            if self.apply_flatfield and not self.flatfields:
                self.get_flatfields()
            if self.use_registration and self.placement != "global":
                self.calculate_shifts(self.timepoints[0], self.regions[0])
            from concurrent.futures import ThreadPoolExecutor
            from .shard import wells_for_rank
            band_mode = self.world > 1 and (self.split_regions or len(self.regions) < self.world)
            my_regions = list(self.regions) if band_mode else \
                [self.regions[i] for i in wells_for_rank(len(self.regions), self.world, self.rank)]
            work = [(t, r) for t in self.timepoints for r in my_regions]
            for timepoint in self.timepoints:
                os.makedirs(os.path.join(self.output_folder, f"{timepoint}_stitched"), exist_ok=True)
            # three stages in flight: decode of region i+1 (thread pool), GPU fusion of region i (this thread),
            # OME-Zarr write of region i-1 (writer thread)
            from .pipeline import RegionPipeline
            fast = (os.environ.get("SB_NO_FAST_IO") is None and self.placement != "global" and
                    RegionPipeline.eligible(self.dtype, _ffi.BLEND_MODES[self.blend_mode], self.output_format, self.chunks))
            pipe = RegionPipeline(self.ctx, (self.input_height, self.input_width), self.chunks[-2:]) if fast and work else None
            self.fast_io_used = pipe is not None
            self.band_mode = self.world > 1 and (self.split_regions or len(self.regions) < self.world)
            if self.band_mode and pipe is None:
                raise RuntimeError("splitting a region over several GPUs needs the OME-Zarr fast path (uint16, paste mode, "
                                   "lattice placement, power-of-two chunk width)")
            if self.band_mode:
                with ThreadPoolExecutor(max_workers=1) as writer:
                    last_path = self._run_bands(pipe, writer)
                work = []
            with ThreadPoolExecutor(max_workers=1) as prefetch, ThreadPoolExecutor(max_workers=1) as writer:
                nxt = prefetch.submit(self._decode_region, *work[0]) if work else None
                pending_write = None
                lane_writes = {}                                   # lane -> write still reading that lane's pinned buffers
                for i, (timepoint, region) in enumerate(work):
                    self.check_stop()
                    loaded = nxt.result()
                    nxt = prefetch.submit(self._decode_region, *work[i + 1]) if i + 1 < len(work) else None
                    if pipe is None:
                        stitched = self.stitch_region(timepoint, region, loaded=loaded)
                        if pending_write is not None:
                            last_path = pending_write.result()
                        self.emit_status(f"Saving... (Timepoint:{timepoint} Region:{region})", is_saving=True)
                        pending_write = writer.submit(self.save_region_ome_zarr, timepoint, region, stitched,
                                                      self.num_pyramid_levels)  # per-region value, fixed before the next region
                        continue
                    # fast path: pinned staging in, level 0 back in zarr-chunk order + device-side pyramid, nothing re-tiled
                    start = time.time()
                    job, cshape, n_data = self._region_job(timepoint, region, loaded)
                    if self.apply_flatfield:
                        self._sync_fields()
                    busy = lane_writes.pop(pipe.next, None)
                    if busy is not None:
                        last_path = busy.result()              # the lane's pinned buffers are free again
                    n_levels = self.num_pyramid_levels
                    ticket = pipe.submit(job, cshape, n_levels, self.apply_flatfield)
                    self.emit_progress(n_data, n_data)
                    print(f"(Timepoint:{timepoint}, Region:{region}) Stitching enqueued in {time.time() - start:.1f}s\n")
                    self.emit_status(f"Saving... (Timepoint:{timepoint} Region:{region})", is_saving=True)
                    lane_writes[ticket["lane"]] = writer.submit(self._save_ticket, timepoint, region, pipe, ticket)
                    pending_write = lane_writes[ticket["lane"]]
                for fut in lane_writes.values():
                    last_path = fut.result()
                if pipe is None and pending_write is not None:
                    last_path = pending_write.result()
            if pipe is not None:
                pipe.close()
            self.check_stop()
            self.emit_complete(last_path, self.dtype)
            print(f"Processing complete. Total time: {time.time() - stime:.1f}s")
        except Exception as exc:
            print(f"Error in StitcherProcess: {exc}")
            if self.status_queue is not None:
                self.status_queue.put(("error", str(exc)))
            raise
        finally:
            if self._ctx is not None:
                self._ctx.close()
                self._ctx = None
