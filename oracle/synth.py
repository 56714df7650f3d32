"""Synthetic Squid acquisitions with known ground-truth tile offsets (CPU oracle side).

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.

Recipe: a per-region "world" image = gaussian-filtered
power-law noise scaled to the uint16 range; tiles are cut on a lattice of step
``round((1-overlap)*W)`` plus integer stage jitter (constant per column for x,
per row for y, so the reference's 2-shift lattice model can represent it) and get
independent per-tile sensor noise.  The on-disk layout written by
:func:`write_squid_layout` is the one ``parse_acquisition_metadata`` consumes
(``stitcher_process.py:261-371``).
"""
from __future__ import annotations

import json
import os
from typing import Dict, List, Sequence, Tuple

import numpy as np
from scipy import ndimage

from .stitch_ref import RegionState, TileRec

# 8.0 um binned sensor pixel, 16x objective on a 180 mm tube lens -> exactly 0.5 um / px
ACQ_PARAMS = {
    "objective": {"magnification": 16.0, "tube_lens_f_mm": 180.0, "name": "16x"},
    "sensor_pixel_size_um": 8.0,
    "tube_lens_mm": 180.0,
    "pixel_binning": 2,
    "dz(um)": 1.5,
}


def pixel_size_um(acq=ACQ_PARAMS) -> float:
    """stitcher_process.py:249-258."""
    obj_focal = acq["objective"]["tube_lens_f_mm"] / acq["objective"]["magnification"]
    return acq["sensor_pixel_size_um"] / (acq["tube_lens_mm"] / obj_focal)


def make_world(h: int, w: int, rng: np.random.Generator, sigma: float = 1.0, power: float = 6.0,
               amp: float = 40000.0, floor: float = 400.0) -> np.ndarray:
    """Smooth blob texture with a heavy tail: enough structure for phase correlation."""
    u = rng.random((h, w), dtype=np.float32) ** power
    g = ndimage.gaussian_filter(u, sigma)
    g = g / g.max()
    coarse = ndimage.gaussian_filter(rng.random((h, w), dtype=np.float32), 12.0)
    coarse = (coarse - coarse.min()) / (np.ptp(coarse) + 1e-12)
    return (floor + amp * (0.75 * g + 0.25 * coarse * g.mean() * 8)).astype(np.float32)


def vignette(h: int, w: int, strength: float = 0.35, shift=(0.04, -0.03)) -> np.ndarray:
    """Smooth radial flatfield, float32, mean exactly-ish 1 (BaSiC-like, ``flatfield`` a13)."""
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    r2 = ((yy / (h - 1) - 0.5 - shift[0]) ** 2 + (xx / (w - 1) - 0.5 - shift[1]) ** 2) / 0.5
    ff = 1.0 - strength * r2
    ff = ff / ff.mean()
    return ff.astype(np.float32)


def make_region(rows: int, cols: int, tile_h: int, tile_w: int, *, channels: Sequence[str] = ("Fluorescence 488 nm Ex",),
                num_z: int = 1, overlap: float = 0.10, jitter: int = 3, seed: int = 0, region: str = "A1",
                noise: float = 30.0, use_registration: bool = False, apply_flatfield: bool = False,
                scan_pattern: str = "Unidirectional", registration_channel: str = "",
                flat_strength: float = 0.35) -> Tuple[RegionState, List[TileRec], Dict]:
    """Build one region in memory.  Returns ``(state, tiles_in_paste_order, truth)``.

    ``truth['col_dx'][c]`` / ``truth['row_dy'][r]`` are the integer stage offsets of
    column c / row r from the ideal lattice (what registration should recover as
    deviations of ``h_shift``/``v_shift`` from ``-overlap``).
    """
    rng = np.random.default_rng(seed)
    px = pixel_size_um()
    step_x = int(round((1.0 - overlap) * tile_w))
    step_y = int(round((1.0 - overlap) * tile_h))
    # registration only sees inter-tile *differences*: keep the jitter a lattice
    # (per-column dx drift, per-row dy drift) that the reference's model can express
    col_jx = rng.integers(-jitter, jitter + 1) if jitter else 0     # extra x step per column
    row_jy = rng.integers(-jitter, jitter + 1) if jitter else 0     # extra y step per row
    col_jy = rng.integers(-jitter, jitter + 1) if jitter else 0     # y drift per column
    row_jx = rng.integers(-jitter, jitter + 1) if jitter else 0     # x drift per row
    pad = 4 * jitter * max(rows, cols) + 8
    world_h = tile_h + (rows - 1) * step_y + 2 * pad
    world_w = tile_w + (cols - 1) * step_x + 2 * pad
    worlds = {ch: make_world(world_h, world_w, rng) for ch in channels}

    channels = list(channels)
    flatfields = {}
    if apply_flatfield:
        for ci in range(len(channels)):
            flatfields[ci] = vignette(tile_h, tile_w, flat_strength, shift=(0.04 * (ci + 1), -0.03 * (ci + 1)))

    x0_mm, y0_mm = 10.0, 20.0
    tiles: List[TileRec] = []
    true_px = {}
    for r in range(rows):
        for c in range(cols):
            fov = r * cols + c
            # nominal stage position (what coordinates.csv reports): ideal lattice
            x_mm = x0_mm + c * step_x * px / 1000.0
            y_mm = y0_mm + r * step_y * px / 1000.0
            # true position in the world: lattice + drift
            wx = pad + c * (step_x + col_jx) + r * row_jx + 2 * jitter * max(rows, cols)
            wy = pad + r * (step_y + row_jy) + c * col_jy + 2 * jitter * max(rows, cols)
            true_px[(r, c)] = (int(wx), int(wy))
            for z in range(num_z):
                for ch in channels:
                    img = worlds[ch][wy:wy + tile_h, wx:wx + tile_w]
                    if apply_flatfield:
                        img = img * flatfields[channels.index(ch)]
                    img = img * (1.0 - 0.08 * z) + rng.normal(0.0, noise, img.shape).astype(np.float32)
                    pixels = np.clip(img, 0, 65535).astype(np.uint16)
                    name = f"{region}_{fov}_{z}_{ch.replace(' ', '_')}.tiff"
                    tiles.append(TileRec(x_mm=x_mm, y_mm=y_mm, z_level=z, channel=ch, pixels=pixels,
                                         fov=fov, name=name))
    tiles.sort(key=lambda t: t.name)     # reference paste order: sorted file names (283-288)

    st = RegionState(tile_h=tile_h, tile_w=tile_w, pixel_size_um=px, pixel_binning=ACQ_PARAMS["pixel_binning"],
                     monochrome_channels=sorted(channels), channel_names=sorted(channels), num_z=num_z,
                     use_registration=use_registration, apply_flatfield=apply_flatfield,
                     scan_pattern=scan_pattern, registration_channel=registration_channel,
                     flatfields={sorted(channels).index(channels[ci]): ff for ci, ff in flatfields.items()})
    truth = {
        "step": (step_y, step_x),
        "h_shift": (int(col_jy), int(-(tile_w - step_x) + col_jx)),   # expected (dy, dx) of calculate_horizontal_shift
        "v_shift": (int(-(tile_h - step_y) + row_jy), int(row_jx)),
        "true_px": true_px,
    }
    return st, tiles, truth


def write_squid_layout(root: str, regions: Dict[str, List[TileRec]], timepoint: int = 0, acq=ACQ_PARAMS) -> None:
    """Materialise regions in the Squid folder layout."""
    import cv2
    os.makedirs(os.path.join(root, str(timepoint)), exist_ok=True)
    with open(os.path.join(root, "acquisition parameters.json"), "w") as f:
        json.dump(acq, f)
    rows = ["region,fov,z_level,x (mm),y (mm),z (um)"]
    seen = set()
    for region, tiles in regions.items():
        for t in tiles:
            key = (region, t.fov, t.z_level)
            if key not in seen:
                seen.add(key)
                rows.append(f"{region},{t.fov},{t.z_level},{t.x_mm!r},{t.y_mm!r},{t.z_level * acq.get('dz(um)', 1.0)!r}")
            px = t.pixels[:, :, ::-1] if (t.pixels.ndim == 3 and t.pixels.shape[2] == 3) else t.pixels   # OpenCV stores BGR
            ok = cv2.imwrite(os.path.join(root, str(timepoint), t.name), np.ascontiguousarray(px))
            if not ok:
                raise IOError(f"cv2.imwrite failed for {t.name}")
    with open(os.path.join(root, str(timepoint), "coordinates.csv"), "w") as f:
        f.write("\n".join(rows) + "\n")
