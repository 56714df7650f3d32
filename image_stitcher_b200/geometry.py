"""Host-side placement geometry of the fusion path (product code; pure integer/float64 Python).

The CUDA kernels receive integers only; everything the reference derives from stage
coordinates and the registration lattice is computed here with the reference's own
rounding rules so that the canvas is bit-identical:

* canvas size          -- ``calculate_output_dimensions`` (stitcher_process.py:423-477)
* tile origin          -- ``stitch_region`` (stitcher_process.py:919-942)
* seam crops           -- ``place_single_channel_tile`` (stitcher_process.py:789-806)
* strip widths         -- ``calculate_shifts`` (stitcher_process.py:602-609)

Quirks kept on purpose: the registered canvas height uses
``H - v_shift[0]`` (over-allocates for the usual negative v_shift); stage coordinates
are truncated with ``int()``; Python ``//`` floors negative numbers.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np


@dataclass(frozen=True)
class Lattice:
    """Registration result applied to every region (self.h_shift / v_shift / h_shift_rev*)."""
    h_shift: Tuple[int, int] = (0, 0)
    v_shift: Tuple[int, int] = (0, 0)
    h_shift_rev: Tuple[int, int] = (0, 0)
    h_shift_rev_odd: int = 0
    s_pattern: bool = False

    def h_for_row(self, row: int) -> Tuple[int, int]:
        if self.s_pattern and row % 2 == self.h_shift_rev_odd:
            return self.h_shift_rev
        return self.h_shift


@dataclass(frozen=True)
class Placement:
    """Where one tile lands: uncropped origin + seam crops (ints handed to ``sb_tile``)."""
    x: int
    y: int
    crop_t: int = 0
    crop_b: int = 0
    crop_l: int = 0
    crop_r: int = 0


def strip_overlaps(tile_w: int, tile_h: int, x_positions: Sequence[float], y_positions: Sequence[float],
                   pixel_size_um: float, pixel_binning: int) -> Tuple[int, int]:
    """Strip extents used for registration (stitcher_process.py:602-609).

    ``round(|W - dx_px| * 1.05) // 2 * binning`` -- floor-divide first, then scale.
    """
    def extent(n_px: int, pos: Sequence[float]) -> int:
        if len(pos) < 2:          # single row / column: no pair along this axis (the reference raises IndexError here)
            return 0
        step_px = (pos[1] - pos[0]) * 1000 / pixel_size_um
        return round(abs(n_px - step_px) * 1.05) // 2 * pixel_binning
    return extent(tile_w, x_positions), extent(tile_h, y_positions)


def canvas_size(tile_w: int, tile_h: int, x_positions: Sequence[float], y_positions: Sequence[float],
                pixel_size_um: float, lattice: Optional[Lattice]) -> Tuple[int, int]:
    """(width, height) of the stitched region (stitcher_process.py:441-466)."""
    if lattice is not None:
        n_cols, n_rows = len(x_positions), len(y_positions)
        if lattice.s_pattern:
            mh0 = max(abs(lattice.h_shift[0]), abs(lattice.h_shift_rev[0]))
            mh1 = max(abs(lattice.h_shift[1]), abs(lattice.h_shift_rev[1]))
        else:
            mh0, mh1 = abs(lattice.h_shift[0]), abs(lattice.h_shift[1])
        width = int(tile_w + (n_cols - 1) * (tile_w - mh1)) + abs((n_rows - 1) * lattice.v_shift[1])
        # reference quirk (:457): minus a (normally negative) v_shift[0] -> taller than needed
        height = int(tile_h + (n_rows - 1) * (tile_h - lattice.v_shift[0])) + abs((n_cols - 1) * mh0)
        return width, height
    width_mm = max(x_positions) - min(x_positions) + (tile_w * pixel_size_um / 1000)
    height_mm = max(y_positions) - min(y_positions) + (tile_h * pixel_size_um / 1000)
    return int(np.ceil(width_mm * 1000 / pixel_size_um)), int(np.ceil(height_mm * 1000 / pixel_size_um))


def pyramid_levels(width: int, height: int, n_regions: int = 1, region_grid_max_dim: int = 1) -> int:
    """stitcher_process.py:468-475."""
    max_dimension = region_grid_max_dim if n_regions > 1 else 1
    return max(1, math.ceil(np.log2(max(width, height) / 1024 * max_dimension)))


def place_tile(x_mm: float, y_mm: float, tile_w: int, tile_h: int, x_positions: Sequence[float],
               y_positions: Sequence[float], pixel_size_um: float, lattice: Optional[Lattice]) -> Placement:
    """Origin and seam crops of one tile (stitcher_process.py:919-942 and :789-806)."""
    if lattice is None:
        return Placement(int((x_mm - min(x_positions)) * 1000 / pixel_size_um),
                         int((y_mm - min(y_positions)) * 1000 / pixel_size_um))
    col, row = x_positions.index(x_mm), y_positions.index(y_mm)
    n_cols, n_rows = len(x_positions), len(y_positions)
    h, v = lattice.h_for_row(row), lattice.v_shift
    x = int(col * (tile_w + h[1]))
    y = int(row * (tile_h + v[0]))
    y += int((n_cols - 1 - col) * abs(h[0])) if h[0] < 0 else int(col * h[0])
    x += int((n_rows - 1 - row) * abs(v[1])) if v[1] < 0 else int(row * v[1])
    vert = max(0, (-v[0] // 2) - abs(h[0]) // 2)
    horz = max(0, (-h[1] // 2) - abs(v[1]) // 2)
    return Placement(x, y,
                     crop_t=vert if row > 0 else 0, crop_b=vert if row < n_rows - 1 else 0,
                     crop_l=horz if col > 0 else 0, crop_r=horz if col < n_cols - 1 else 0)


def center_pairs(x_positions: Sequence[float], y_positions: Sequence[float], s_pattern: bool):
    """The 2 (3 for S-Pattern) tile pairs ``calculate_shifts`` registers (stitcher_process.py:612-660).

    Returns a list of ``(kind, (x, y) of ref, (x, y) of mov)`` with kind in {'h', 'v', 'h_rev'} and
    ``h_shift_rev_odd``.
    """
    cx, cy = (len(x_positions) - 1) // 2, (len(y_positions) - 1) // 2
    pairs = []
    right = x_positions[cx + 1] if cx + 1 < len(x_positions) else None
    below = y_positions[cy + 1] if cy + 1 < len(y_positions) else None
    if right is not None:
        pairs.append(("h", (x_positions[cx], y_positions[cy]), (right, y_positions[cy])))
    if below is not None:
        pairs.append(("v", (x_positions[cx], y_positions[cy]), (x_positions[cx], below)))
    # truthiness test on the coordinates, as the reference writes it (:649)
    if s_pattern and right and below:
        pairs.append(("h_rev", (x_positions[cx], below), (right, below)))
    return pairs, (cy % 2 == 0)


def grid_pairs(n_rows: int, n_cols: int) -> List[Tuple[str, Tuple[int, int], Tuple[int, int]]]:
    """All adjacent pairs of a rows x cols grid (all-pairs mode behind ``dynamic_registration``)."""
    out = []
    for r in range(n_rows):
        for c in range(n_cols):
            if c + 1 < n_cols:
                out.append(("h", (r, c), (r, c + 1)))
            if r + 1 < n_rows:
                out.append(("v", (r, c), (r + 1, c)))
    return out


def solve_positions(n_tiles: int, tile_w: int, tile_h: int, pairs: Sequence[Tuple[str, int, int]],
                    shifts: Sequence[Tuple[int, int]], weights: Optional[Sequence[float]] = None) -> List[Tuple[int, int]]:
    """Global placement from pairwise shifts (extension; the reference applies ONE lattice to every tile).

    ``pairs[k] = (kind, a, b)`` with kind 'h' (b is the right neighbour of a) or 'v' (b below a) and
    ``shifts[k] = (dy, dx)`` exactly as ``calculate_horizontal_shift`` / ``calculate_vertical_shift`` return them, i.e.
    the origin of b relative to a is ``(dy, tile_w + dx)`` for 'h' and ``(tile_h + dy, dx)`` for 'v'.  Solves the
    least-squares system ``p_b - p_a = d_ab`` (tile 0 anchored, optional per-pair weights) and returns integer pixel
    origins ``(x, y)`` shifted so that the minimum is 0.  With consistent pair shifts the solution is exact.
    """
    m = len(pairs)
    if n_tiles == 0:
        return []
    A = np.zeros((m + 1, n_tiles), np.float64)
    bx = np.zeros(m + 1, np.float64)
    by = np.zeros(m + 1, np.float64)
    for k, ((kind, a, b), (dy, dx)) in enumerate(zip(pairs, shifts)):
        w = 1.0 if weights is None else float(weights[k])
        A[k, a], A[k, b] = -w, w
        if kind == "v":
            bx[k], by[k] = w * dx, w * (tile_h + dy)
        else:
            bx[k], by[k] = w * (tile_w + dx), w * dy
    A[m, 0] = 1.0                                   # anchor
    x = np.linalg.lstsq(A, bx, rcond=None)[0]
    y = np.linalg.lstsq(A, by, rcond=None)[0]
    xi = np.rint(x - x.min()).astype(np.int64)
    yi = np.rint(y - y.min()).astype(np.int64)
    return [(int(a), int(b)) for a, b in zip(xi, yi)]


def _rect_subtract(a, b):
    """``a`` minus ``b`` (``(x0, y0, x1, y1)``, exclusive ends) as up to four disjoint rectangles."""
    ix0, iy0, ix1, iy1 = max(a[0], b[0]), max(a[1], b[1]), min(a[2], b[2]), min(a[3], b[3])
    if ix0 >= ix1 or iy0 >= iy1:
        return [a]
    out = []
    if a[1] < iy0:
        out.append((a[0], a[1], a[2], iy0))
    if iy1 < a[3]:
        out.append((a[0], iy1, a[2], a[3]))
    if a[0] < ix0:
        out.append((a[0], iy0, ix0, iy1))
    if ix1 < a[2]:
        out.append((ix1, iy0, a[2], iy1))
    return out


def visible_boxes(tiles: Sequence[tuple], tile_h: int, tile_w: int, canvas_h: int, canvas_w: int, align: int = 8):
    """Per tile of a PASTE job (``(px, x, y, c, z, crop_t, crop_b, crop_l, crop_r)`` in paste order): the bounding box
    ``(x0, y0, x1, y1)`` in TILE coordinates of the pixels that reach the canvas -- its kept rectangle minus what every
    later tile of the same plane overwrites (stitcher_process.py:817) -- or ``None`` when nothing of it is visible.
    ``x0`` / ``x1`` are rounded outwards to multiples of ``align`` pixels (the paste kernel loads whole 16-byte vectors).
    The fusion kernels never read a pixel outside these boxes, so a host pipeline need not upload the rest."""
    rects = []
    for (_, x, y, c, z, ct, cb, cl, cr) in tiles:
        rects.append((max(x + cl, 0), max(y + ct, 0), min(x + tile_w - cr, canvas_w), min(y + tile_h - cb, canvas_h)))
    out = []
    for i, t in enumerate(tiles):
        r = rects[i]
        if r[0] >= r[2] or r[1] >= r[3]:
            out.append(None)
            continue
        pieces = [r]
        for j in range(i + 1, len(tiles)):
            if tiles[j][3] != t[3] or tiles[j][4] != t[4]:
                continue
            nxt = []
            for pc in pieces:
                nxt += _rect_subtract(pc, rects[j])
            pieces = nxt
            if not pieces:
                break
        if not pieces:
            out.append(None)
            continue
        x0 = min(pc[0] for pc in pieces) - t[1]
        y0 = min(pc[1] for pc in pieces) - t[2]
        x1 = max(pc[2] for pc in pieces) - t[1]
        y1 = max(pc[3] for pc in pieces) - t[2]
        x0 = (x0 // align) * align
        x1 = min(-(-x1 // align) * align, tile_w)
        out.append((max(x0, 0), max(y0, 0), x1, min(y1, tile_h)))
    return out
