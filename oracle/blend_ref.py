"""Definition (not restatement) of the EXTENSION modes: weighted blending and darkfield.

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.

The reference has none of this: it fuses by crop-to-seam + overwrite
(``stitcher_process.py:789-817``) and fits BaSiC without a darkfield (``:521``).
``north_star`` asks for "flatfield-corrected weighted blending", so the behaviour
is *defined here* in float64 NumPy and the CUDA kernels are graded against this
definition (fused uint16 within 1 LSB):

    corrected_i = clip((tile_i - dark) / flat, 0, 65535)        (float, NOT truncated)
    out(p)      = clip(rint(sum_i w_i(p) corrected_i(p) / sum_i w_i(p)), 0, 65535)
                  over every tile i whose cropped rectangle contains p; 0 where none does

    e_x(p) = min(p.x - rx0, rx1 - 1 - p.x) + 1     (1 on the tile's first/last column)
    feather: w = e_x * e_y                          (distance-to-edge tent)
    linear : w = min(e_x, ov_x + 1) * min(e_y, ov_y + 1)   (ramp across the nominal overlap,
                                                            flat in the tile interior)

``rint`` is round-half-to-even.  Tiles are tuples
``(pixels, x, y, c, z, crop_t, crop_b, crop_l, crop_r)`` as passed to ``sb_fuse_region``.
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence

import numpy as np


def correct(tile: np.ndarray, flat: Optional[np.ndarray], dark: Optional[np.ndarray]) -> np.ndarray:
    v = tile.astype(np.float64)
    if dark is not None:
        v = v - dark.astype(np.float64)
    if flat is not None:
        with np.errstate(divide="ignore", invalid="ignore"):
            v = v / flat.astype(np.float64)
    v = np.where(np.isnan(v), 0.0, v)
    return np.clip(v, 0.0, 65535.0)


def edge_weight(n: int, lo_crop: int, hi_crop: int, mode: str, ov: int) -> np.ndarray:
    """Per-index weight along one axis of an n-long tile with crops; zero outside the kept part."""
    idx = np.arange(n, dtype=np.float64)
    lo, hi = lo_crop, n - hi_crop
    e = np.minimum(idx - lo, hi - 1 - idx) + 1
    if mode == "linear":
        e = np.minimum(e, ov + 1)
    elif mode != "feather":
        raise ValueError(mode)
    e[(idx < lo) | (idx >= hi)] = 0
    return e


def fuse_blend(tiles: Sequence[tuple], canvas_shape, mode: str, ov=(0, 0),
               flats: Optional[Dict[int, np.ndarray]] = None, darks: Optional[Dict[int, np.ndarray]] = None):
    """Returns the ``(1, C, Z, H, W)`` uint16 canvas for ``mode`` in {'linear', 'feather'}."""
    C, Z, Hc, Wc = canvas_shape
    num = np.zeros((C, Z, Hc, Wc), np.float64)
    den = np.zeros((C, Z, Hc, Wc), np.float64)
    for (px, x, y, c, z, ct, cb, cl, cr) in tiles:
        h, w = px.shape
        v = correct(px, (flats or {}).get(c), (darks or {}).get(c))
        wy = edge_weight(h, ct, cb, mode, ov[1])
        wx = edge_weight(w, cl, cr, mode, ov[0])
        wt = wy[:, None] * wx[None, :]
        y0, x0 = max(y, 0), max(x, 0)
        y1, x1 = min(y + h, Hc), min(x + w, Wc)
        if y1 <= y0 or x1 <= x0:
            continue
        sl = (slice(y0 - y, y1 - y), slice(x0 - x, x1 - x))
        num[c, z, y0:y1, x0:x1] += wt[sl] * v[sl]
        den[c, z, y0:y1, x0:x1] += wt[sl]
    with np.errstate(divide="ignore", invalid="ignore"):
        out = np.where(den > 0, np.rint(num / den), 0.0)
    return np.clip(out, 0, 65535).astype(np.uint16)[None]


def fuse_paste(tiles: Sequence[tuple], canvas_shape, flats=None, darks=None, field_dtype=np.float32):
    """Paste mode with the (extension) darkfield: trunc(clip((tile - dark) / flat)) in ``field_dtype``,
    later tiles overwrite earlier ones.  Without a darkfield this equals the reference's
    ``place_single_channel_tile`` (stitcher_process.py:771-826)."""
    C, Z, Hc, Wc = canvas_shape
    out = np.zeros((1, C, Z, Hc, Wc), np.uint16)
    for (px, x, y, c, z, ct, cb, cl, cr) in tiles:
        v = px
        flat, dark = (flats or {}).get(c), (darks or {}).get(c)
        if flat is not None or dark is not None:
            f = px.astype(field_dtype)
            if dark is not None:
                f = f - dark.astype(field_dtype)
            if flat is not None:
                with np.errstate(divide="ignore", invalid="ignore"):
                    f = f / flat.astype(field_dtype)
            f = np.where(np.isnan(f), 0, f)
            v = np.clip(f, 0, 65535).astype(np.uint16)
        h, w = px.shape
        v = v[ct:h - cb, cl:w - cr]
        xx, yy = x + cl, y + ct
        y1, x1 = min(yy + v.shape[0], Hc), min(xx + v.shape[1], Wc)
        out[0, c, z, yy:y1, xx:x1] = v[:y1 - yy, :x1 - xx]
    return out
