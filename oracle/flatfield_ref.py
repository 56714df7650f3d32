"""TEST INFRASTRUCTURE ONLY -- CPU definition of the flat-field ESTIMATOR extension.

The reference fits its flat-fields with BaSiC (``get_flatfields``, stitcher_process.py:505-571: up to 32 random tiles
per timepoint, stop above 48, ``BaSiC(get_darkfield=False, smoothness_flatfield=1).fit(images)``; the result feeds
``apply_flatfield_correction`` :828-842 as an ``H x W`` float array with mean about 1).  BaSiCPy (and its pinned JAX)
are third-party, absent from ``/root/reference`` and from this image, and the reference holds no fitted field or test
for them, so BaSiC itself cannot be restated or pinned here: **parity unpinned, own definition**.  What the product
offers when BaSiCPy is not importable is the robust estimator below, and this file is its oracle:

1. every tile is reduced to a ``g x g`` grid of cell means (cell of pixel ``(y, x)`` = ``(y*g // H, x*g // W)``; integer sums);
2. each tile's grid is divided by the tile's mean intensity (all-zero tiles are dropped);
3. per cell, the MEDIAN over the tiles (sparse bright foreground does not move a median the way it moves a mean);
4. separable Gaussian smoothing of the grid (sigma in cells, radius ceil(3 sigma), symmetric boundary ``d c b a | a b c d``);
5. division by the grid mean (so the field has mean 1 like BaSiC's);
6. bilinear interpolation from the cell centres to ``H x W``, cast to float32.
"""
from __future__ import annotations

import math

import numpy as np


def cell_edges(n: int, g: int) -> np.ndarray:
    """First pixel of every cell (and ``n`` at the end): pixel ``p`` belongs to cell ``p * g // n``."""
    return np.array([-(-c * n // g) for c in range(g)] + [n], dtype=np.int64)


def cell_sums(tiles: np.ndarray, g: int) -> np.ndarray:
    """``(n, g, g)`` integer sums of the pixels of every cell."""
    n, h, w = tiles.shape
    ey, ex = cell_edges(h, g), cell_edges(w, g)
    t64 = tiles.astype(np.int64)
    rows = np.add.reduceat(t64, ey[:-1], axis=1)
    return np.add.reduceat(rows, ex[:-1], axis=2)


def gaussian_weights(sigma: float) -> np.ndarray:
    r = int(math.ceil(3.0 * sigma))
    k = np.arange(-r, r + 1, dtype=np.float64)
    w = np.exp(-(k * k) / (2.0 * sigma * sigma)) if sigma > 0 else np.ones(1)
    return w / w.sum()


def smooth_axis(a: np.ndarray, weights: np.ndarray, axis: int) -> np.ndarray:
    r = (len(weights) - 1) // 2
    n = a.shape[axis]
    out = np.zeros_like(a)
    idx = np.arange(n)
    for k in range(-r, r + 1):
        j = idx + k
        # symmetric boundary (d c b a | a b c d), folded as often as needed for short axes
        period = 2 * n
        j = np.mod(j, period)
        j = np.where(j >= n, period - 1 - j, j)
        out += weights[k + r] * np.take(a, j, axis=axis)
    return out


def estimate_flatfield(tiles: np.ndarray, grid: int = 128, sigma: float = 2.0) -> np.ndarray:
    """``tiles``: ``(n, H, W)`` uint8 / uint16.  Returns the ``H x W`` float32 field (all ones without a usable tile)."""
    tiles = np.asarray(tiles)
    n, h, w = tiles.shape
    g = max(1, min(int(grid), h, w))
    sums = cell_sums(tiles, g).astype(np.float64)
    ey, ex = cell_edges(h, g), cell_edges(w, g)
    count = (np.diff(ey)[:, None] * np.diff(ex)[None, :]).astype(np.float64)
    tile_mean = sums.sum(axis=(1, 2)) / float(h * w)
    keep = tile_mean > 0
    if not keep.any():
        return np.ones((h, w), np.float32)
    v = (sums[keep] / count[None]) / tile_mean[keep][:, None, None]
    med = np.median(v, axis=0)
    wts = gaussian_weights(sigma)
    sm = smooth_axis(smooth_axis(med, wts, 1), wts, 0)
    mean = sm.mean()
    f = sm / mean if mean > 0 else np.ones_like(sm)
    # bilinear interpolation from the cell centres
    def axis_coords(n_px):
        u = (np.arange(n_px, dtype=np.float64) + 0.5) * g / n_px - 0.5
        u = np.clip(u, 0.0, g - 1.0)
        i0 = np.minimum(np.floor(u).astype(np.int64), max(g - 2, 0))
        return i0, u - i0
    j0, fy = axis_coords(h)
    i0, fx = axis_coords(w)
    j1, i1 = np.minimum(j0 + 1, g - 1), np.minimum(i0 + 1, g - 1)
    top = f[j0][:, i0] * (1.0 - fx)[None, :] + f[j0][:, i1] * fx[None, :]
    bot = f[j1][:, i0] * (1.0 - fx)[None, :] + f[j1][:, i1] * fx[None, :]
    return (top * (1.0 - fy)[:, None] + bot * fy[:, None]).astype(np.float32)
