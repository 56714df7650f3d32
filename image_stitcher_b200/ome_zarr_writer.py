"""Minimal OME-Zarr (NGFF 0.4 on zarr v2) writer for stitched regions.

The reference writes its canvases through third-party stacks (ome_zarr / bioio / aicsimageio,
stitcher_process.py:958-1549) none of which is installed here, and north_star keeps those writers
"as they are".  This module exists so that BASELINE.json configs[0] ("... stitch to OME-Zarr") runs
end to end without them.  It reproduces the *metadata* the reference emits in
``save_region_ome_zarr`` (stitcher_process.py:1039-1124):

* axes t, c, z, y, x with units second / micrometer (:1085-1091);
* one ``scale`` transform per level ``[1, 1, dz, px * 2**l, px * 2**l]`` (:1066-1078);
* ``omero.channels`` with label, 6-hex colour, window 0..dtype max (:1099-1118);
* chunks ``(1, 1, 1, 2048, 2048)`` (stitcher_process.py:161) or ``(1, 1, 1, 512, 512)`` (stitcher.py:235);
* the pyramid is nearest-neighbour x2 per level (``Scaler.nearest``, :1061-1062) -- here ``[::2, ::2]``.

Differences, stated: chunks are stored uncompressed by default (``compressor: null``; the reference's
Blosc default is not available offline) or with the zarr-v2 ``zlib`` codec; nested ``/`` chunk keys.

Two entry points: ``write_ome_zarr`` takes the row-major ``(1, C, Z, Hc, Wc)`` canvas the reference's
writers take; ``write_ome_zarr_chunked`` takes the library's SB_LAYOUT_CHUNKED output (zarr chunk
order, edge chunks zero-padded) and writes every chunk with one contiguous ``tofile``.
"""
from __future__ import annotations

import json
import os
import zlib
from typing import Optional, Sequence

import numpy as np

AXES = [
    {"name": "t", "type": "time", "unit": "second"},
    {"name": "c", "type": "channel"},
    {"name": "z", "type": "space", "unit": "micrometer"},
    {"name": "y", "type": "space", "unit": "micrometer"},
    {"name": "x", "type": "space", "unit": "micrometer"},
]


def _dump(path: str, obj) -> None:
    with open(path, "w") as fh:
        json.dump(obj, fh, indent=2)


def _zarray(shape, chunks, dtype: np.dtype, compressor: Optional[str]):
    return {
        "zarr_format": 2,
        "shape": [int(s) for s in shape],
        "chunks": [int(c) for c in chunks],
        "dtype": np.dtype(dtype).newbyteorder("<").str if np.dtype(dtype).itemsize > 1 else np.dtype(dtype).str,
        "compressor": {"id": "zlib", "level": 1} if compressor == "zlib" else None,
        "fill_value": 0,
        "order": "C",
        "filters": None,
        "dimension_separator": "/",
    }


def _encode(chunk: np.ndarray, compressor: Optional[str]) -> bytes:
    raw = np.ascontiguousarray(chunk).tobytes()
    return zlib.compress(raw, 1) if compressor == "zlib" else raw


def _write_chunk(level_dir: str, idx: Sequence[int], chunk: np.ndarray, compressor: Optional[str]) -> None:
    d = os.path.join(level_dir, *[str(i) for i in idx[:-1]])
    os.makedirs(d, exist_ok=True)
    path = os.path.join(d, str(idx[-1]))
    if compressor is None and chunk.flags.c_contiguous:
        chunk.tofile(path)
    else:
        with open(path, "wb") as fh:
            fh.write(_encode(chunk, compressor))


def _write_level(level_dir: str, data: np.ndarray, chunks, compressor: Optional[str]) -> None:
    """``data`` is (T, C, Z, H, W); edge chunks are padded with the fill value (zarr v2 semantics)."""
    os.makedirs(level_dir, exist_ok=True)
    ch_y, ch_x = int(chunks[3]), int(chunks[4])
    _dump(os.path.join(level_dir, ".zarray"), _zarray(data.shape, (1, 1, 1, ch_y, ch_x), data.dtype, compressor))
    T, C, Z, H, W = data.shape
    for t in range(T):
        for c in range(C):
            for z in range(Z):
                plane = data[t, c, z]
                for iy in range(-(-H // ch_y)):
                    for ix in range(-(-W // ch_x)):
                        blk = plane[iy * ch_y:(iy + 1) * ch_y, ix * ch_x:(ix + 1) * ch_x]
                        if blk.shape != (ch_y, ch_x):
                            full = np.zeros((ch_y, ch_x), dtype=data.dtype)
                            full[:blk.shape[0], :blk.shape[1]] = blk
                            blk = full
                        _write_chunk(level_dir, (t, c, z, iy, ix), np.ascontiguousarray(blk), compressor)


def _group_attrs(name, n_levels, pixel_size_um, dz_um, channel_names, channel_colors, dtype):
    datasets = [{"path": str(l),
                 "coordinateTransformations": [{"type": "scale",
                                                "scale": [1, 1, float(dz_um), float(pixel_size_um * 2 ** l),
                                                          float(pixel_size_um * 2 ** l)]}]}
                for l in range(n_levels)]
    vmax = int(np.iinfo(dtype).max) if np.issubdtype(dtype, np.integer) else 1
    return {
        "multiscales": [{"version": "0.4", "name": name, "axes": AXES, "datasets": datasets}],
        "omero": {"id": 1, "name": name, "version": "0.4",
                  "channels": [{"label": str(nm), "color": f"{int(col):06X}",
                                "window": {"start": 0, "end": vmax, "min": 0, "max": vmax},
                                "active": True, "coefficient": 1, "family": "linear"}
                               for nm, col in zip(channel_names, channel_colors)]},
    }


def write_ome_zarr(path: str, data: np.ndarray, *, pixel_size_um: float, dz_um: float = 1.0,
                   channel_names: Sequence[str], channel_colors: Sequence[int], num_levels: int = 1,
                   chunks=(1, 1, 1, 2048, 2048), compressor: Optional[str] = None, name: Optional[str] = None,
                   levels: Optional[Sequence[np.ndarray]] = None) -> str:
    """Write a (1, C, Z, H, W) canvas as a multiscale OME-Zarr group; returns ``path``.  ``levels`` takes levels
    ``1 ..`` already made on the GPU (``sb_pyramid``); missing ones are sliced on the host."""
    if data.ndim != 5:
        raise ValueError(f"expected a 5-D TCZYX array, got shape {data.shape}")
    if compressor not in (None, "zlib"):
        raise ValueError("compressor must be None or 'zlib'")
    os.makedirs(path, exist_ok=True)
    _dump(os.path.join(path, ".zgroup"), {"zarr_format": 2})
    level = data
    n_written = 0
    for l in range(max(1, int(num_levels))):
        if l > 0:
            if levels is not None and l - 1 < len(levels):
                level = levels[l - 1]
            else:
                level = level[..., ::2, ::2]                   # nearest-neighbour x2 (Scaler.nearest)
            if level.shape[-1] < 1 or level.shape[-2] < 1:
                break
        _write_level(os.path.join(path, str(l)), level, chunks, compressor)
        n_written += 1
    _dump(os.path.join(path, ".zattrs"),
          _group_attrs(name or os.path.basename(path).replace(".ome.zarr", ""), n_written, pixel_size_um, dz_um,
                       channel_names, channel_colors, data.dtype))
    return path


def write_ome_zarr_chunked(path: str, chunked: np.ndarray, shape, chunk_hw, *, pixel_size_um: float, dz_um: float = 1.0,
                           channel_names: Sequence[str], channel_colors: Sequence[int], name: Optional[str] = None,
                           levels: Optional[Sequence[np.ndarray]] = None) -> str:
    """Level 0 from the library's SB_LAYOUT_CHUNKED buffer: ``chunked`` is ``(C * Z, ncy, ncx, chunk_h, chunk_w)`` -- each
    chunk contiguous, edge chunks already zero-padded -- and ``shape`` the logical ``(C, Z, Hc, Wc)``.  No re-tiling pass
    on the host: one write per chunk.  ``levels`` = the multiscale levels 1.. as ``(1, C, Z, h, w)`` arrays
    (``sb_pyramid``); they are small and go through the ordinary per-level writer."""
    C, Z, H, W = (int(v) for v in shape)
    ch_y, ch_x = int(chunk_hw[0]), int(chunk_hw[1])
    ncy, ncx = -(-H // ch_y), -(-W // ch_x)
    buf = np.asarray(chunked).reshape(C * Z, ncy, ncx, ch_y, ch_x)
    os.makedirs(os.path.join(path, "0"), exist_ok=True)
    _dump(os.path.join(path, ".zgroup"), {"zarr_format": 2})
    _dump(os.path.join(path, "0", ".zarray"), _zarray((1, C, Z, H, W), (1, 1, 1, ch_y, ch_x), buf.dtype, None))
    for c in range(C):
        for z in range(Z):
            for iy in range(ncy):
                for ix in range(ncx):
                    _write_chunk(os.path.join(path, "0"), (0, c, z, iy, ix), buf[c * Z + z, iy, ix], None)
    n_written = 1
    for l, level in enumerate(levels or [], start=1):
        _write_level(os.path.join(path, str(l)), np.asarray(level), (1, 1, 1, ch_y, ch_x), None)
        n_written += 1
    _dump(os.path.join(path, ".zattrs"),
          _group_attrs(name or os.path.basename(path).replace(".ome.zarr", ""), n_written, pixel_size_um, dz_um,
                       channel_names, channel_colors, buf.dtype))
    return path


def pyramid_level_shapes(height: int, width: int, n_levels: int):
    """``(h_l, w_l)`` of levels ``0 .. n_levels - 1`` under ``[..., ::2, ::2]`` per level."""
    out = [(int(height), int(width))]
    for _ in range(1, max(1, int(n_levels))):
        out.append(((out[-1][0] + 1) // 2, (out[-1][1] + 1) // 2))
    return out


def _dump_atomic(path: str, obj) -> None:
    tmp = f"{path}.{os.getpid()}.tmp"
    _dump(tmp, obj)
    os.replace(tmp, path)                                  # several workers write the same metadata: never a torn file


def _chunk_view(level_dir: str, idx: Sequence[int], chunk_hw, dtype) -> np.memmap:
    """Read-write view of one uncompressed chunk file, created at its full size if it does not exist yet (a sparse file
    reads as the fill value 0).  Workers that own different rows of the same chunk write through their own views."""
    d = os.path.join(level_dir, *[str(i) for i in idx[:-1]])
    os.makedirs(d, exist_ok=True)
    path = os.path.join(d, str(idx[-1]))
    nbytes = int(chunk_hw[0]) * int(chunk_hw[1]) * np.dtype(dtype).itemsize
    fd = os.open(path, os.O_RDWR | os.O_CREAT, 0o644)
    try:
        if os.fstat(fd).st_size < nbytes:
            os.ftruncate(fd, nbytes)
    finally:
        os.close(fd)
    return np.memmap(path, dtype=dtype, mode="r+", shape=(int(chunk_hw[0]), int(chunk_hw[1])))


def write_ome_zarr_band(path: str, chunked: np.ndarray, levels: Sequence[np.ndarray], *, plane, row0: int, full_shape,
                        chunk_hw, n_levels: int, pixel_size_um: float, dz_um: float = 1.0, channel_names: Sequence[str],
                        channel_colors: Sequence[int], name: Optional[str] = None) -> str:
    """One worker's share of a region that several GPUs fuse together (the reference's out-of-core
    analogue is zarr_stitcher.py:570-612): the rows ``[row0, row0 + band_h)`` of plane ``(c, z)``.

    ``chunked`` is the band's level 0 in the library's chunk order ``(1, ncy_band, ncx, chunk_h, chunk_w)`` (``row0`` is a
    multiple of ``chunk_h``, so the band owns whole chunks: one ``tofile`` each); ``levels`` are the band's multiscale
    levels ``1 ..`` as ``(1, 1, 1, h, w)`` arrays -- ``row0`` is a multiple of every ``2**l``, so decimating the band
    equals decimating the region -- whose rows land INSIDE chunks shared with other workers: written through a memory map
    of the (uncompressed, full-size) chunk file.  Every worker writes the same metadata (atomically)."""
    C, Z, H, W = (int(v) for v in full_shape)
    c, z = int(plane[0]), int(plane[1])
    ch_y, ch_x = int(chunk_hw[0]), int(chunk_hw[1])
    if row0 % ch_y:
        raise ValueError(f"band origin {row0} is not a multiple of the chunk height {ch_y}")
    buf = np.asarray(chunked)
    ncx = -(-W // ch_x)
    buf = buf.reshape(-1, ncx, ch_y, ch_x)
    shapes = pyramid_level_shapes(H, W, n_levels)
    n_written = 1 + min(len(levels), len(shapes) - 1)
    os.makedirs(path, exist_ok=True)
    _dump_atomic(os.path.join(path, ".zgroup"), {"zarr_format": 2})
    for l in range(n_written):
        os.makedirs(os.path.join(path, str(l)), exist_ok=True)
        _dump_atomic(os.path.join(path, str(l), ".zarray"),
                     _zarray((1, C, Z, shapes[l][0], shapes[l][1]), (1, 1, 1, ch_y, ch_x), buf.dtype, None))
    _dump_atomic(os.path.join(path, ".zattrs"),
                 _group_attrs(name or os.path.basename(path).replace(".ome.zarr", ""), n_written, pixel_size_um, dz_um,
                              channel_names, channel_colors, buf.dtype))
    for iy in range(buf.shape[0]):
        for ix in range(ncx):
            _write_chunk(os.path.join(path, "0"), (0, c, z, row0 // ch_y + iy, ix), buf[iy, ix], None)
    for l in range(1, n_written):
        lv = np.asarray(levels[l - 1])
        lv = lv.reshape(lv.shape[-2], lv.shape[-1])
        r0 = row0 >> l
        if lv.shape[1] != shapes[l][1] or r0 + lv.shape[0] > shapes[l][0]:
            raise ValueError(f"level {l}: band {lv.shape} at row {r0} does not fit the region level {shapes[l]}")
        y = r0
        while y < r0 + lv.shape[0]:
            iy = y // ch_y
            y1 = min((iy + 1) * ch_y, r0 + lv.shape[0])
            for ix in range(-(-lv.shape[1] // ch_x)):
                x0, x1 = ix * ch_x, min((ix + 1) * ch_x, lv.shape[1])
                mm = _chunk_view(os.path.join(path, str(l)), (0, c, z, iy, ix), (ch_y, ch_x), buf.dtype)
                mm[y - iy * ch_y:y1 - iy * ch_y, :x1 - x0] = lv[y - r0:y1 - r0, x0:x1]
                mm.flush()
                del mm
            y = y1
    return path


def read_ome_zarr_level(path: str, level: int = 0) -> np.ndarray:
    """Read one level back into a dense array (used by the tests; zarr itself is not installed)."""
    ldir = os.path.join(path, str(level))
    with open(os.path.join(ldir, ".zarray")) as fh:
        za = json.load(fh)
    shape, chunks = za["shape"], za["chunks"]
    dtype = np.dtype(za["dtype"])
    out = np.zeros(shape, dtype=dtype)
    grid = [-(-s // c) for s, c in zip(shape, chunks)]
    for idx in np.ndindex(*grid):
        p = os.path.join(ldir, *[str(i) for i in idx])
        if not os.path.exists(p):
            continue
        with open(p, "rb") as fh:
            raw = fh.read()
        if za["compressor"]:
            raw = zlib.decompress(raw)
        blk = np.frombuffer(raw, dtype=dtype).reshape(chunks)
        sl = tuple(slice(i * c, min((i + 1) * c, s)) for i, c, s in zip(idx, chunks, shape))
        out[sl] = blk[tuple(slice(0, s.stop - s.start) for s in sl)]
    return out
