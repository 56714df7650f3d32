"""NumPy restatement of the reference's registration + fusion methods (CPU oracle).

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.

Every function cites the lines of ``/root/reference/stitcher_process.py`` it
follows (``stitcher.py`` holds AST-identical copies).  The
reference works on files + dask arrays; here a region is an in-memory list of
:class:`TileRec` in the reference's paste order (sorted file names,
``stitcher_process.py:283-288``) so the arithmetic can be exercised without the
third-party readers/writers.  ``tests/golden`` pins this file against the
unmodified reference run under ``oracle/ref_shim.py``.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from .pcc_ref import phase_cross_correlation


@dataclass
class TileRec:
    """One image file of a region: what ``acquisition_metadata[key]`` holds (309-319)."""
    x_mm: float
    y_mm: float
    z_level: int
    channel: str
    pixels: np.ndarray          # H x W (mono), H x W x 3 (RGB) or 1 x H x W
    fov: int = 0
    name: str = ""              # file name; region order == sorted(name)


@dataclass
class RegionState:
    """The attributes of ``StitcherProcess`` the hot path reads (146-168)."""
    tile_h: int
    tile_w: int
    pixel_size_um: float
    pixel_binning: int = 1
    dtype: np.dtype = np.dtype(np.uint16)
    monochrome_channels: List[str] = field(default_factory=list)
    channel_names: List[str] = field(default_factory=list)
    num_z: int = 1
    use_registration: bool = False
    apply_flatfield: bool = False
    scan_pattern: str = "Unidirectional"
    registration_channel: str = ""
    registration_z_level: int = 0
    flatfields: Dict[int, np.ndarray] = field(default_factory=dict)
    h_shift: Tuple[int, int] = (0, 0)
    v_shift: Tuple[int, int] = (0, 0)
    h_shift_rev: Tuple[int, int] = (0, 0)
    h_shift_rev_odd: int = 0
    n_regions: int = 1
    region_grid_max_dim: int = 1

    @property
    def num_c(self) -> int:
        return len(self.monochrome_channels)


# --------------------------------------------------------------------------- registration

def normalize_image(img: np.ndarray, dtype=np.uint16) -> np.ndarray:
    """stitcher_process.py:844-855 -- whole-tile min/max stretch, float64, truncating cast."""
    img = np.asarray(img)
    img_min, img_max = img.min(), img.max()
    with np.errstate(invalid="ignore", divide="ignore"):
        img_normalized = (img - img_min) / (img_max - img_min)
        scale_factor = np.iinfo(dtype).max if np.issubdtype(dtype, np.integer) else 1
        scaled = img_normalized * scale_factor
        if img_max == img_min:
            # 0/0 -> NaN -> undefined cast in the reference; x86 yields 0.  The
            # build defines this case as all-zero (hard part 3).
            return np.zeros(img.shape, dtype=dtype)
        return scaled.astype(dtype)


def horizontal_strips(img_left, img_right, max_overlap):
    """stitcher_process.py:677-679 (inputs already normalised)."""
    margin = int(img_left.shape[0] * 0.25)
    a = img_left[margin:-margin, -max_overlap:]
    b = img_right[margin:-margin, :max_overlap]
    return a, b


def vertical_strips(img_top, img_bot, max_overlap):
    """stitcher_process.py:700-702 (inputs already normalised)."""
    margin = int(img_top.shape[1] * 0.25)
    a = img_top[-max_overlap:, margin:-margin]
    b = img_bot[:max_overlap, margin:-margin]
    return a, b


def calculate_horizontal_shift(img_left, img_right, max_overlap, dtype=np.uint16,
                               upsample_factor=10, return_details=False):
    """stitcher_process.py:664-685.  Returns ``(dy, dx)`` Python ints (half-even ``round``)."""
    a, b = horizontal_strips(normalize_image(img_left, dtype), normalize_image(img_right, dtype), max_overlap)
    res = phase_cross_correlation(a, b, upsample_factor=upsample_factor, return_details=return_details)
    shift = res[0]
    out = (round(shift[0]), round(shift[1] - a.shape[1]))
    return (out, shift, res[3]) if return_details else out


def calculate_vertical_shift(img_top, img_bot, max_overlap, dtype=np.uint16,
                             upsample_factor=10, return_details=False):
    """stitcher_process.py:687-708."""
    a, b = vertical_strips(normalize_image(img_top, dtype), normalize_image(img_bot, dtype), max_overlap)
    res = phase_cross_correlation(a, b, upsample_factor=upsample_factor, return_details=return_details)
    shift = res[0]
    out = (round(shift[0] - a.shape[0]), round(shift[1]))
    return (out, shift, res[3]) if return_details else out


def strip_overlaps(st: RegionState, x_pos_list: Sequence[float], y_pos_list: Sequence[float]):
    """stitcher_process.py:602-609 -- note ``round(..) // 2 * binning`` precedence."""
    dx_mm = x_pos_list[1] - x_pos_list[0]
    dy_mm = y_pos_list[1] - y_pos_list[0]
    dx_pixels = dx_mm * 1000 / st.pixel_size_um
    dy_pixels = dy_mm * 1000 / st.pixel_size_um
    max_x_overlap = round(abs(st.tile_w - dx_pixels) * 1.05) // 2 * st.pixel_binning
    max_y_overlap = round(abs(st.tile_h - dy_pixels) * 1.05) // 2 * st.pixel_binning
    return max_x_overlap, max_y_overlap


def get_tile(tiles: Sequence[TileRec], x, y, channel, z_level) -> Optional[np.ndarray]:
    """stitcher_process.py:710-737 -- first exact (x, y, channel, z) match in region order."""
    for t in tiles:
        if t.x_mm == x and t.y_mm == y and t.channel == channel and t.z_level == z_level:
            return t.pixels
    return None


def calculate_shifts(st: RegionState, tiles: Sequence[TileRec], upsample_factor=10) -> RegionState:
    """stitcher_process.py:573-662.  Mutates and returns ``st`` (h_shift, v_shift, h_shift_rev*)."""
    x_positions = sorted(set(t.x_mm for t in tiles))
    y_positions = sorted(set(t.y_mm for t in tiles))
    st.h_shift = (0, 0)
    st.v_shift = (0, 0)
    if not st.registration_channel or st.registration_channel not in st.channel_names:
        st.registration_channel = st.channel_names[0]                      # 589-594

    max_x_overlap, max_y_overlap = strip_overlaps(st, x_positions, y_positions)
    cx = (len(x_positions) - 1) // 2                                        # 613-614
    cy = (len(y_positions) - 1) // 2
    center_x, center_y = x_positions[cx], y_positions[cy]
    right_x = bottom_y = None
    ch, z = st.registration_channel, st.registration_z_level

    if cx + 1 < len(x_positions):                                           # 623-633
        right_x = x_positions[cx + 1]
        a = get_tile(tiles, center_x, center_y, ch, z)
        b = get_tile(tiles, right_x, center_y, ch, z)
        if a is not None and b is not None:
            st.h_shift = calculate_horizontal_shift(a, b, max_x_overlap, st.dtype, upsample_factor)
    if cy + 1 < len(y_positions):                                           # 636-646
        bottom_y = y_positions[cy + 1]
        a = get_tile(tiles, center_x, center_y, ch, z)
        b = get_tile(tiles, center_x, bottom_y, ch, z)
        if a is not None and b is not None:
            st.v_shift = calculate_vertical_shift(a, b, max_y_overlap, st.dtype, upsample_factor)
    if st.scan_pattern == "S-Pattern" and right_x and bottom_y:             # 649-660
        a = get_tile(tiles, center_x, bottom_y, ch, z)
        b = get_tile(tiles, right_x, bottom_y, ch, z)
        if a is not None and b is not None:
            st.h_shift_rev = calculate_horizontal_shift(a, b, max_x_overlap, st.dtype, upsample_factor)
            st.h_shift_rev_odd = cy % 2 == 0
    return st


# --------------------------------------------------------------------------- fusion geometry

def output_dimensions(st: RegionState, x_positions: Sequence[float], y_positions: Sequence[float]):
    """stitcher_process.py:441-477.  Returns ``(width, height, num_pyramid_levels)``.

    Reproduces the canvas-height quirk at :457 (``H - v_shift[0]`` with a negative
    ``v_shift[0]`` over-allocates).
    """
    if st.use_registration:
        num_cols, num_rows = len(x_positions), len(y_positions)
        if st.scan_pattern == "S-Pattern":
            max_h = (max(abs(st.h_shift[0]), abs(st.h_shift_rev[0])),
                     max(abs(st.h_shift[1]), abs(st.h_shift_rev[1])))
        else:
            max_h = (abs(st.h_shift[0]), abs(st.h_shift[1]))
        width = int(st.tile_w + ((num_cols - 1) * (st.tile_w - max_h[1])))
        width += abs((num_rows - 1) * st.v_shift[1])
        height = int(st.tile_h + ((num_rows - 1) * (st.tile_h - st.v_shift[0])))
        height += abs((num_cols - 1) * max_h[0])
    else:
        width_mm = max(x_positions) - min(x_positions) + (st.tile_w * st.pixel_size_um / 1000)
        height_mm = max(y_positions) - min(y_positions) + (st.tile_h * st.pixel_size_um / 1000)
        width = int(np.ceil(width_mm * 1000 / st.pixel_size_um))
        height = int(np.ceil(height_mm * 1000 / st.pixel_size_um))
    max_dimension = st.region_grid_max_dim if st.n_regions > 1 else 1
    levels = max(1, math.ceil(np.log2(max(width, height) / 1024 * max_dimension)))
    return width, height, levels


def tile_position(st: RegionState, t: TileRec, x_positions, y_positions):
    """stitcher_process.py:919-942.  Returns ``(x_pixel, y_pixel, col_index, row_index)``."""
    if st.use_registration:
        col = x_positions.index(t.x_mm)
        row = y_positions.index(t.y_mm)
        if st.scan_pattern == "S-Pattern" and row % 2 == st.h_shift_rev_odd:
            h = st.h_shift_rev
        else:
            h = st.h_shift
        x_pixel = int(col * (st.tile_w + h[1]))
        y_pixel = int(row * (st.tile_h + st.v_shift[0]))
        if h[0] < 0:
            y_pixel += int((len(x_positions) - 1 - col) * abs(h[0]))
        else:
            y_pixel += int(col * h[0])
        if st.v_shift[1] < 0:
            x_pixel += int((len(y_positions) - 1 - row) * abs(st.v_shift[1]))
        else:
            x_pixel += int(row * st.v_shift[1])
        return x_pixel, y_pixel, col, row
    x_min, y_min = min(x_positions), min(y_positions)
    x_pixel = int((t.x_mm - x_min) * 1000 / st.pixel_size_um)
    y_pixel = int((t.y_mm - y_min) * 1000 / st.pixel_size_um)
    return x_pixel, y_pixel, None, None


def seam_crops(st: RegionState, col, row, n_cols, n_rows):
    """stitcher_process.py:789-799.  Returns ``(top, bottom, left, right)``; zeros if unregistered."""
    if not st.use_registration:
        return 0, 0, 0, 0
    if st.scan_pattern == "S-Pattern" and row % 2 == st.h_shift_rev_odd:
        h = st.h_shift_rev
    else:
        h = st.h_shift
    v = st.v_shift
    top = max(0, (-v[0] // 2) - abs(h[0]) // 2) if row > 0 else 0
    bottom = max(0, (-v[0] // 2) - abs(h[0]) // 2) if row < n_rows - 1 else 0
    left = max(0, (-h[1] // 2) - abs(v[1]) // 2) if col > 0 else 0
    right = max(0, (-h[1] // 2) - abs(v[1]) // 2) if col < n_cols - 1 else 0
    return top, bottom, left, right


def apply_flatfield_correction(st: RegionState, tile: np.ndarray, channel_idx: int) -> np.ndarray:
    """stitcher_process.py:828-842 -- divide, clip to the dtype range, truncating cast."""
    if channel_idx in st.flatfields:
        with np.errstate(invalid="ignore", divide="ignore"):
            q = (tile / st.flatfields[channel_idx]).clip(min=np.iinfo(st.dtype).min,
                                                         max=np.iinfo(st.dtype).max)
            # NaN (0/0) has an undefined cast in the reference; x86 yields 0.
            q = np.where(np.isnan(q), 0, q)
        return q.astype(st.dtype)
    return tile


def place_single_channel_tile(st, canvas, tile, x_pixel, y_pixel, z_level, channel_idx,
                              col, row, n_cols, n_rows):
    """stitcher_process.py:771-826 -- flatfield, seam crop, clip to canvas, overwrite."""
    if st.apply_flatfield:
        tile = apply_flatfield_correction(st, tile, channel_idx)
    top, bottom, left, right = seam_crops(st, col, row, n_cols, n_rows)
    if st.use_registration:
        tile = tile[top:tile.shape[0] - bottom, left:tile.shape[1] - right]
        x_pixel += left
        y_pixel += top
    y_end = min(y_pixel + tile.shape[0], canvas.shape[3])
    x_end = min(x_pixel + tile.shape[1], canvas.shape[4])
    tile_slice = tile[:y_end - y_pixel, :x_end - x_pixel]
    canvas[0, channel_idx, z_level, y_pixel:y_end, x_pixel:x_end] = tile_slice


def place_tile(st, canvas, t: TileRec, x_pixel, y_pixel, col, row, n_cols, n_rows):
    """stitcher_process.py:739-769 -- mono / RGB / 1xHxW dispatch; always timepoint 0."""
    px = t.pixels
    if px.ndim == 2:
        c = st.monochrome_channels.index(t.channel)
        place_single_channel_tile(st, canvas, px, x_pixel, y_pixel, t.z_level, c, col, row, n_cols, n_rows)
    elif px.ndim == 3:
        if px.shape[2] == 3:
            base = t.channel.split("_")[0]
            for i, color in enumerate(["R", "G", "B"]):
                c = st.monochrome_channels.index(f"{base}_{color}")
                place_single_channel_tile(st, canvas, px[:, :, i], x_pixel, y_pixel, t.z_level, c,
                                          col, row, n_cols, n_rows)
        elif px.shape[0] == 1:
            c = st.monochrome_channels.index(t.channel)
            place_single_channel_tile(st, canvas, px[0], x_pixel, y_pixel, t.z_level, c, col, row, n_cols, n_rows)
    else:
        raise ValueError(f"Unexpected tile shape: {px.shape}")


def stitch_region(st: RegionState, tiles: Sequence[TileRec]) -> np.ndarray:
    """stitcher_process.py:883-956 + init_output 489-503.  ``tiles`` in paste order.

    Returns the ``(1, C, Z, Hc, Wc)`` canvas (what ``.compute()`` at :1993-1994 yields).
    """
    x_positions = sorted(set(t.x_mm for t in tiles))
    y_positions = sorted(set(t.y_mm for t in tiles))
    width, height, _ = output_dimensions(st, x_positions, y_positions)
    canvas = np.zeros((1, st.num_c, st.num_z, height, width), dtype=st.dtype)
    for t in tiles:
        x_pixel, y_pixel, col, row = tile_position(st, t, x_positions, y_positions)
        place_tile(st, canvas, t, x_pixel, y_pixel, col, row, len(x_positions), len(y_positions))
    return canvas
