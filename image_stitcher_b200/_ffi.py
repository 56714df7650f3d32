"""ctypes binding of libstitchb200.so (include/stitchb200.h).

The library is the product; this file only marshals pointers.  There is no CPU
fallback: if the shared library is missing or no CUDA device is usable the
import-time loader / ``Context()`` raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SB_LIB_PATH") or os.path.join(_HERE, "_lib", "libstitchb200.so")

SB_MEM_HOST, SB_MEM_DEVICE = 0, 1
SB_U16, SB_U8 = 0, 1
SB_FIELD_F32, SB_FIELD_F64 = 0, 1
SB_BLEND_PASTE, SB_BLEND_LINEAR, SB_BLEND_FEATHER = 0, 1, 2
SB_LAYOUT_ROWMAJOR, SB_LAYOUT_CHUNKED = 0, 1
SB_PREC_F32, SB_PREC_F64, SB_PREC_AUTO = 0, 1, 2
SB_DIR_HORIZONTAL, SB_DIR_VERTICAL = 0, 1
BLEND_MODES = {"paste": SB_BLEND_PASTE, "linear": SB_BLEND_LINEAR, "feather": SB_BLEND_FEATHER}

EXPORTS = [
    "sb_version", "sb_create", "sb_destroy", "sb_last_error", "sb_kernel_launches", "sb_num_lanes",
    "sb_device_sm_count", "sb_host_alloc", "sb_host_free", "sb_device_alloc", "sb_device_free",
    "sb_memcpy_h2d", "sb_memcpy_d2h", "sb_memcpy_async", "sb_memcpy2d_async", "sb_set_flatfield", "sb_set_darkfield", "sb_clear_fields",
    "sb_flatfield_apply", "sb_fuse_region", "sb_fuse_regions", "sb_sync", "sb_lane_mark", "sb_lane_wait_mark", "sb_set_lane_stream", "sb_canvas_pitch",
    "sb_chunked_plane_elems", "sb_register_pairs", "sb_register_pairs_async", "sb_normalize",
    "sb_pyramid_elems", "sb_pyramid", "sb_estimate_flatfield", "sb_selftest", "sb_debug_read", "sb_debug_tc_profile",
]


class SbTile(C.Structure):
    _fields_ = [("px", C.c_void_p), ("x", C.c_int32), ("y", C.c_int32), ("c", C.c_int32), ("z", C.c_int32),
                ("crop_t", C.c_int32), ("crop_b", C.c_int32), ("crop_l", C.c_int32), ("crop_r", C.c_int32)]


class SbFuseJob(C.Structure):
    _fields_ = [("tiles", C.POINTER(SbTile)), ("n_tiles", C.c_int32), ("tile_h", C.c_int32), ("tile_w", C.c_int32),
                ("dtype", C.c_int32), ("tile_mem", C.c_int32), ("num_c", C.c_int32), ("num_z", C.c_int32),
                ("height", C.c_int32), ("width", C.c_int32), ("apply_flatfield", C.c_int32), ("blend", C.c_int32),
                ("blend_ov_x", C.c_int32), ("blend_ov_y", C.c_int32), ("out", C.c_void_p), ("out_mem", C.c_int32),
                ("out_layout", C.c_int32), ("out_row_pitch", C.c_int64), ("chunk_h", C.c_int32),
                ("chunk_w", C.c_int32), ("field_c0", C.c_int32), ("reserved0", C.c_int32)]


class SbPair(C.Structure):
    _fields_ = [("ref", C.c_void_p), ("mov", C.c_void_p), ("dir", C.c_int32), ("reserved", C.c_int32)]


class SbPairResult(C.Structure):
    _fields_ = [("dy", C.c_int32), ("dx", C.c_int32), ("shift", C.c_double * 2), ("coarse", C.c_int32 * 2),
                ("fine", C.c_int32 * 2), ("peak", C.c_float), ("second", C.c_float), ("runner_up", C.c_float), ("fine_peak", C.c_float),
                ("fine_second", C.c_float),
                ("ref_min", C.c_int32), ("ref_max", C.c_int32), ("mov_min", C.c_int32), ("mov_max", C.c_int32),
                ("precision", C.c_int32)]


class SbRegisterJob(C.Structure):
    _fields_ = [("pairs", C.POINTER(SbPair)), ("n_pairs", C.c_int32), ("tile_h", C.c_int32), ("tile_w", C.c_int32),
                ("dtype", C.c_int32), ("mem", C.c_int32), ("max_overlap_x", C.c_int32), ("max_overlap_y", C.c_int32),
                ("upsample_factor", C.c_int32), ("precision", C.c_int32), ("lane", C.c_int32)]


_lib = None


def load_library(path: Optional[str] = None):
    """dlopen the C-ABI library and declare its prototypes.  Raises if it has not been built."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    path = path or LIB_PATH
    if not os.path.isfile(path):
        raise RuntimeError(f"{path} not found: build it with `python -m image_stitcher_b200.build` "
                           "(there is no CPU fallback for the CUDA hot path)")
    lib = C.CDLL(path)
    vp, i32, i64 = C.c_void_p, C.c_int, C.c_int64
    lib.sb_version.restype = i32
    lib.sb_create.argtypes = [i32, C.POINTER(vp)]
    lib.sb_destroy.argtypes = [vp]
    lib.sb_destroy.restype = None
    lib.sb_last_error.argtypes = [vp]
    lib.sb_last_error.restype = C.c_char_p
    lib.sb_kernel_launches.argtypes = [vp]
    lib.sb_kernel_launches.restype = i64
    lib.sb_num_lanes.argtypes = [vp]
    lib.sb_device_sm_count.argtypes = [vp]
    lib.sb_host_alloc.argtypes = [vp, C.c_size_t]
    lib.sb_host_alloc.restype = vp
    lib.sb_host_free.argtypes = [vp, vp]
    lib.sb_host_free.restype = None
    lib.sb_device_alloc.argtypes = [vp, C.c_size_t]
    lib.sb_device_alloc.restype = vp
    lib.sb_device_free.argtypes = [vp, vp]
    lib.sb_device_free.restype = None
    lib.sb_memcpy_h2d.argtypes = [vp, vp, vp, C.c_size_t]
    lib.sb_memcpy_d2h.argtypes = [vp, vp, vp, C.c_size_t]
    lib.sb_memcpy_async.argtypes = [vp, i32, vp, vp, C.c_size_t, i32]
    lib.sb_memcpy2d_async.argtypes = [vp, i32, vp, C.c_size_t, vp, C.c_size_t, C.c_size_t, C.c_size_t, i32]
    for fn in (lib.sb_set_flatfield, lib.sb_set_darkfield):
        fn.argtypes = [vp, i32, vp, i32, i32, i32, i32]
    lib.sb_clear_fields.argtypes = [vp]
    lib.sb_flatfield_apply.argtypes = [vp, i32, vp, vp, i32, i32, i32, i32, i32]
    lib.sb_fuse_region.argtypes = [vp, C.POINTER(SbFuseJob), i32]
    lib.sb_fuse_regions.argtypes = [vp, C.POINTER(SbFuseJob), i32, i32]
    lib.sb_sync.argtypes = [vp, i32]
    lib.sb_set_lane_stream.argtypes = [vp, i32, vp]
    lib.sb_lane_mark.argtypes = [vp, i32]
    lib.sb_lane_wait_mark.argtypes = [vp, i32, i32]
    lib.sb_canvas_pitch.argtypes = [C.c_int32]
    lib.sb_canvas_pitch.restype = i64
    lib.sb_chunked_plane_elems.argtypes = [C.c_int32] * 4
    lib.sb_chunked_plane_elems.restype = i64
    lib.sb_register_pairs.argtypes = [vp, C.POINTER(SbRegisterJob), C.POINTER(SbPairResult)]
    lib.sb_register_pairs_async.argtypes = [vp, C.POINTER(SbRegisterJob), C.POINTER(SbPairResult)]
    lib.sb_normalize.argtypes = [vp, vp, vp, i32, i32, i32, i32, i32]
    lib.sb_estimate_flatfield.argtypes = [vp, C.POINTER(C.c_void_p), i32, i32, i32, i32, i32, i32, C.c_double, vp, i32]
    lib.sb_pyramid_elems.argtypes = [i32, i32, i32, i32]
    lib.sb_pyramid_elems.restype = i64
    lib.sb_pyramid.argtypes = [vp, vp, i32, i32, i32, i32, i64, i32, i32, vp, i32, i32]
    lib.sb_selftest.argtypes = [vp, i32, i64, C.POINTER(C.c_uint64)]
    lib.sb_debug_read.argtypes = [vp, i32, i32, vp, i64]
    lib.sb_debug_read.restype = i64
    lib.sb_debug_tc_profile.argtypes = [vp, C.POINTER(C.c_longlong)]
    if path == LIB_PATH:
        _lib = lib
    return lib


def _pixel_dtype(a, default=SB_U16) -> int:
    """SB_U8 / SB_U16 of a numpy array or torch tensor; raw addresses take ``default``."""
    dt = str(getattr(a, "dtype", ""))
    if dt.endswith("uint8"):
        return SB_U8
    if dt.endswith("uint16") or dt.endswith("int16"):
        return SB_U16
    if dt == "":
        return default
    raise TypeError(f"unsupported pixel dtype {dt} (uint8 and uint16 are implemented)")


def _ptr(a) -> int:
    """Address of a numpy array's first element, a raw int address, or a torch tensor's data_ptr."""
    if isinstance(a, (int, np.integer)):
        return int(a)
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    if hasattr(a, "data_ptr"):
        return int(a.data_ptr())
    raise TypeError(f"cannot take the address of {type(a)}")


def _check_tile(a, tile_shape, i):
    """The library reads tile_h * tile_w pixels from every tile pointer: an ndarray of another shape, or a
    non-contiguous one, must not be handed over (raw addresses and device tensors are the caller's responsibility)."""
    if isinstance(a, np.ndarray):
        if tuple(a.shape) != (int(tile_shape[0]), int(tile_shape[1])):
            raise ValueError(f"tile {i}: shape {a.shape} != tile_shape {tuple(tile_shape)}")
        if not a.flags.c_contiguous:
            raise ValueError(f"tile {i}: array is not C-contiguous")


class Context:
    """One ``sb_ctx``: create it lazily inside the worker process (after fork), one per device."""

    def __init__(self, device: int = 0):
        self.lib = load_library()
        h = C.c_void_p()
        rc = self.lib.sb_create(int(device), C.byref(h))
        if rc != 0:
            raise RuntimeError(f"sb_create(device={device}) failed [{rc}]: {self.lib.sb_last_error(None).decode()}")
        self.handle = h
        self.device = device
        self._pinned = {}

    # ------------------------------------------------------------------ plumbing
    def _check(self, rc: int, what: str):
        if rc != 0:
            raise RuntimeError(f"{what} failed [{rc}]: {self.lib.sb_last_error(self.handle).decode()}")

    def close(self):
        if getattr(self, "handle", None):
            for addr in list(self._pinned):
                self.lib.sb_host_free(self.handle, addr)
            self._pinned.clear()
            self.lib.sb_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    @property
    def kernel_launches(self) -> int:
        return int(self.lib.sb_kernel_launches(self.handle))

    @property
    def num_lanes(self) -> int:
        return int(self.lib.sb_num_lanes(self.handle))

    @property
    def sm_count(self) -> int:
        return int(self.lib.sb_device_sm_count(self.handle))

    def pinned_empty(self, shape, dtype) -> np.ndarray:
        """A numpy array over page-locked host memory (freed when the context closes)."""
        dtype = np.dtype(dtype)
        n = int(np.prod(shape)) * dtype.itemsize
        addr = self.lib.sb_host_alloc(self.handle, max(n, 1))
        if not addr:
            raise MemoryError(self.lib.sb_last_error(self.handle).decode())
        self._pinned[addr] = n
        buf = (C.c_uint8 * max(n, 1)).from_address(addr)
        return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def sync(self, lane: int = -1):
        self._check(self.lib.sb_sync(self.handle, lane), "sb_sync")

    def lane_mark(self, lane: int):
        self._check(self.lib.sb_lane_mark(self.handle, lane), "sb_lane_mark")

    def lane_wait_mark(self, lane: int, other: int):
        self._check(self.lib.sb_lane_wait_mark(self.handle, lane, other), "sb_lane_wait_mark")

    def set_lane_stream(self, lane: int, stream_ptr: Optional[int]):
        self._check(self.lib.sb_set_lane_stream(self.handle, lane, C.c_void_p(stream_ptr or 0)), "sb_set_lane_stream")

    # ------------------------------------------------------------------ fields
    @staticmethod
    def _field_args(field, mem):
        if mem == SB_MEM_HOST:
            field = np.ascontiguousarray(field)
            if field.dtype not in (np.float32, np.float64):
                field = field.astype(np.float64)
            dt = SB_FIELD_F64 if field.dtype == np.float64 else SB_FIELD_F32
            return field, dt, field.shape
        dt = SB_FIELD_F64 if "float64" in str(field.dtype) else SB_FIELD_F32
        return field, dt, tuple(field.shape)

    def set_flatfield(self, channel: int, field, mem: int = SB_MEM_HOST):
        field, dt, shape = self._field_args(field, mem)
        self._check(self.lib.sb_set_flatfield(self.handle, channel, _ptr(field), dt, mem, shape[0], shape[1]),
                    "sb_set_flatfield")

    def set_darkfield(self, channel: int, field, mem: int = SB_MEM_HOST):
        field, dt, shape = self._field_args(field, mem)
        self._check(self.lib.sb_set_darkfield(self.handle, channel, _ptr(field), dt, mem, shape[0], shape[1]),
                    "sb_set_darkfield")

    def clear_fields(self):
        self._check(self.lib.sb_clear_fields(self.handle), "sb_clear_fields")

    def flatfield_apply(self, channel: int, tiles: np.ndarray) -> np.ndarray:
        tiles = np.ascontiguousarray(tiles)
        t3 = tiles.reshape((-1,) + tiles.shape[-2:])
        out = np.empty_like(t3)
        self._check(self.lib.sb_flatfield_apply(self.handle, channel, _ptr(t3), _ptr(out), t3.shape[0], t3.shape[1],
                                                t3.shape[2], _pixel_dtype(t3), SB_MEM_HOST), "sb_flatfield_apply")
        return out.reshape(tiles.shape)

    # ------------------------------------------------------------------ fusion
    def fuse_region(self, tiles: Sequence[tuple], tile_shape, canvas_shape, *, out, tile_mem=SB_MEM_HOST,
                    out_mem=SB_MEM_HOST, apply_flatfield=False, blend=SB_BLEND_PASTE, blend_ov=(0, 0),
                    layout=SB_LAYOUT_ROWMAJOR, out_row_pitch=0, chunk=(0, 0), lane=-1, keepalive=None, dtype=None,
                    field_c0=0):
        """``tiles``: sequence of ``(px, x, y, c, z, crop_t, crop_b, crop_l, crop_r)`` in paste order.

        ``canvas_shape`` = ``(num_c, num_z, height, width)``.  ``out`` is a numpy array / address / tensor.
        ``field_c0``: the job fuses a window of a region's planes numbered from 0; tile channel ``c`` takes the
        context's field of channel ``field_c0 + c``.
        """
        n = len(tiles)
        arr = (SbTile * max(n, 1))()
        for i, t in enumerate(tiles):
            px, x, y, c, z, ct, cb, cl, cr = t
            _check_tile(px, tile_shape, i)
            arr[i] = SbTile(_ptr(px), int(x), int(y), int(c), int(z), int(ct), int(cb), int(cl), int(cr))
        if dtype is None:                      # pixel dtype: from the first tile (or the canvas) when they are arrays
            dtype = _pixel_dtype(tiles[0][0] if n else out, _pixel_dtype(out))
            if isinstance(out, np.ndarray) and _pixel_dtype(out) != dtype:
                raise TypeError(f"canvas dtype {out.dtype} does not match the tile dtype (the canvas has the tile dtype)")
        job = SbFuseJob(arr, n, int(tile_shape[0]), int(tile_shape[1]), int(dtype), tile_mem,
                        int(canvas_shape[0]), int(canvas_shape[1]), int(canvas_shape[2]), int(canvas_shape[3]),
                        int(bool(apply_flatfield)), int(blend), int(blend_ov[0]), int(blend_ov[1]),
                        _ptr(out), out_mem, layout, int(out_row_pitch), int(chunk[0]), int(chunk[1]), int(field_c0), 0)
        self._check(self.lib.sb_fuse_region(self.handle, C.byref(job), lane), "sb_fuse_region")

    def fuse_regions(self, jobs: Sequence[dict], *, lane=-1):
        """``sb_fuse_regions``: every element of ``jobs`` holds the keyword arguments of :meth:`fuse_region`
        (``tiles``, ``tile_shape``, ``canvas_shape``, ``out`` ...).  Regions of one geometry in device memory are fused by
        one launch; anything else region by region."""
        arrs, cjobs = [], []
        for kw in jobs:
            kw = dict(kw)
            tiles = kw.pop("tiles")
            n = len(tiles)
            arr = (SbTile * max(n, 1))()
            ts, cs, out = kw.pop("tile_shape"), kw.pop("canvas_shape"), kw.pop("out")
            for i, (px, x, y, c, z, ct, cb, cl, cr) in enumerate(tiles):
                _check_tile(px, ts, i)
                arr[i] = SbTile(_ptr(px), int(x), int(y), int(c), int(z), int(ct), int(cb), int(cl), int(cr))
            arrs.append(arr)
            dtype = kw.get("dtype")
            if dtype is None:
                dtype = _pixel_dtype(tiles[0][0] if n else out, _pixel_dtype(out))
            bo, ch = kw.get("blend_ov", (0, 0)), kw.get("chunk", (0, 0))
            cjobs.append(SbFuseJob(arr, n, int(ts[0]), int(ts[1]), int(dtype), kw.get("tile_mem", SB_MEM_HOST), int(cs[0]), int(cs[1]),
                                   int(cs[2]), int(cs[3]), int(bool(kw.get("apply_flatfield", False))),
                                   int(kw.get("blend", SB_BLEND_PASTE)), int(bo[0]), int(bo[1]), _ptr(out),
                                   kw.get("out_mem", SB_MEM_HOST), kw.get("layout", SB_LAYOUT_ROWMAJOR),
                                   int(kw.get("out_row_pitch", 0)), int(ch[0]), int(ch[1]), int(kw.get("field_c0", 0)), 0))
        cj = (SbFuseJob * max(len(cjobs), 1))(*cjobs)
        self._check(self.lib.sb_fuse_regions(self.handle, cj, len(cjobs), lane), "sb_fuse_regions")

    # ------------------------------------------------------------------ registration
    def device_alloc(self, nbytes: int) -> int:
        addr = self.lib.sb_device_alloc(self.handle, int(nbytes))
        if not addr:
            raise MemoryError(self.lib.sb_last_error(self.handle).decode())
        return int(addr)

    def device_free(self, addr: int):
        self.lib.sb_device_free(self.handle, C.c_void_p(addr))

    def memcpy_async(self, lane: int, dst, src, nbytes: int, kind: int):
        """kind: 0 = H2D, 1 = D2H, 2 = D2D, on the lane's stream."""
        self._check(self.lib.sb_memcpy_async(self.handle, lane, _ptr(dst), _ptr(src), int(nbytes), kind),
                    "sb_memcpy_async")

    def memcpy2d_async(self, lane: int, dst, dst_pitch: int, src, src_pitch: int, width_bytes: int, height: int, kind: int):
        """``height`` rows of ``width_bytes`` bytes between pitched buffers (addresses or arrays), on the lane's stream."""
        self._check(self.lib.sb_memcpy2d_async(self.handle, lane, _ptr(dst), int(dst_pitch), _ptr(src), int(src_pitch),
                                               int(width_bytes), int(height), kind), "sb_memcpy2d_async")

    @staticmethod
    def _pair_dicts(res):
        return [{"dy": r.dy, "dx": r.dx, "shift": (r.shift[0], r.shift[1]),
                 "coarse": (r.coarse[0], r.coarse[1]), "fine": (r.fine[0], r.fine[1]), "peak": r.peak,
                 "second": r.second, "runner_up": r.runner_up, "fine_peak": r.fine_peak, "fine_second": r.fine_second, "ref_minmax": (r.ref_min, r.ref_max),
                 "mov_minmax": (r.mov_min, r.mov_max), "precision": r.precision} for r in res]

    def _register_job(self, pairs, tile_shape, max_overlap_x, max_overlap_y, mem, upsample_factor, precision, lane,
                      dtype=None):
        n = len(pairs)
        if dtype is None:
            dtype = _pixel_dtype(pairs[0][0])
        arr = (SbPair * n)()
        for i, (ref, mov, d) in enumerate(pairs):
            _check_tile(ref, tile_shape, i)
            _check_tile(mov, tile_shape, i)
            arr[i] = SbPair(_ptr(ref), _ptr(mov), int(d), 0)
        res = (SbPairResult * n)()
        job = SbRegisterJob(arr, n, int(tile_shape[0]), int(tile_shape[1]), int(dtype), mem, int(max_overlap_x),
                            int(max_overlap_y), int(upsample_factor), int(precision), int(lane))
        return arr, res, job

    def register_pairs(self, pairs: Sequence[tuple], tile_shape, max_overlap_x: int, max_overlap_y: int, *,
                       mem=SB_MEM_HOST, upsample_factor: int = 10, precision: int = SB_PREC_AUTO, lane: int = 0,
                       dtype=None):
        """``pairs``: sequence of ``(ref, mov, dir)``.  Returns a list of dicts (see ``sb_pair_result``)."""
        if len(pairs) == 0:
            return []
        arr, res, job = self._register_job(pairs, tile_shape, max_overlap_x, max_overlap_y, mem, upsample_factor,
                                           precision, lane, dtype)
        self._check(self.lib.sb_register_pairs(self.handle, C.byref(job), res), "sb_register_pairs")
        return self._pair_dicts(res)

    def register_pairs_async(self, pairs: Sequence[tuple], tile_shape, max_overlap_x: int, max_overlap_y: int, *,
                             mem=SB_MEM_HOST, upsample_factor: int = 10, precision: int = SB_PREC_AUTO, lane: int = 0,
                             dtype=None):
        """Enqueue on ``lane`` and return a :class:`PendingRegistration`; its ``get()`` is valid after ``sync(lane)``
        (``sb_register_pairs_async``).  The tiles must stay valid until then."""
        if len(pairs) == 0:
            return PendingRegistration(self, lane, None, None, None, list(pairs))
        arr, res, job = self._register_job(pairs, tile_shape, max_overlap_x, max_overlap_y, mem, upsample_factor,
                                           precision, lane, dtype)
        self._check(self.lib.sb_register_pairs_async(self.handle, C.byref(job), res), "sb_register_pairs_async")
        return PendingRegistration(self, lane, arr, res, job, list(pairs))

    def normalize(self, tiles: np.ndarray) -> np.ndarray:
        tiles = np.ascontiguousarray(tiles)
        t3 = tiles.reshape((-1,) + tiles.shape[-2:])
        out = np.empty_like(t3)
        self._check(self.lib.sb_normalize(self.handle, _ptr(t3), _ptr(out), t3.shape[0], t3.shape[1], t3.shape[2],
                                          _pixel_dtype(t3), SB_MEM_HOST), "sb_normalize")
        return out.reshape(tiles.shape)

    def estimate_flatfield(self, tiles, *, grid: int = 128, sigma: float = 2.0, mem=SB_MEM_HOST, tile_shape=None,
                           dtype=None) -> np.ndarray:
        """Robust flat-field estimate (``sb_estimate_flatfield``, an extension -- not BaSiC) from a sequence of
        same-shaped 2-D tiles (numpy arrays, or device addresses with ``mem=SB_MEM_DEVICE`` and ``tile_shape``)."""
        tiles = list(tiles)
        if not tiles:
            raise ValueError("no tiles")
        if mem == SB_MEM_HOST:
            tiles = [np.ascontiguousarray(t) for t in tiles]
            tile_shape = tiles[0].shape
            if any(t.shape != tile_shape or t.dtype != tiles[0].dtype for t in tiles) or len(tile_shape) != 2:
                raise ValueError("tiles must be 2-D arrays of one shape and dtype")
        if dtype is None:
            dtype = _pixel_dtype(tiles[0])
        ptrs = (C.c_void_p * len(tiles))(*[_ptr(t) for t in tiles])
        out = np.empty((int(tile_shape[0]), int(tile_shape[1])), dtype=np.float32)
        self._check(self.lib.sb_estimate_flatfield(self.handle, ptrs, len(tiles), int(tile_shape[0]), int(tile_shape[1]),
                                                   int(dtype), mem, int(grid), float(sigma), _ptr(out), SB_MEM_HOST),
                    "sb_estimate_flatfield")
        return out

    def debug_read(self, lane: int, which: int, max_elems: int) -> np.ndarray:
        """Test hook (``sb_debug_read``): complex64 intermediates of the lane's last float32 registration group."""
        out = np.empty(int(max_elems), np.complex64)
        n = int(self.lib.sb_debug_read(self.handle, lane, which, _ptr(out), out.nbytes))
        if n < 0:
            raise RuntimeError(f"sb_debug_read failed [{n}]: {self.lib.sb_last_error(self.handle).decode()}")
        return out[:n // 8]

    def selftest(self, which: int, arg: int):
        """Test hook (``sb_selftest``): returns ``(cases_checked, mismatches, first_mismatch_key | max_error, flag)``."""
        out = (C.c_uint64 * 4)()
        self._check(self.lib.sb_selftest(self.handle, int(which), int(arg), out), "sb_selftest")
        return int(out[0]), int(out[1]), int(out[2]), int(out[3])

    @staticmethod
    def pyramid_shapes(canvas_shape, n_levels: int):
        """``(planes, h_l, w_l)`` of levels ``1 .. n_levels - 1`` (``[..., ::2, ::2]`` per level)."""
        planes, h, w = int(np.prod(canvas_shape[:-2], dtype=np.int64)), int(canvas_shape[-2]), int(canvas_shape[-1])
        shapes = []
        for _ in range(1, max(1, int(n_levels))):
            h, w = (h + 1) // 2, (w + 1) // 2
            shapes.append((planes, h, w))
        return shapes

    def pyramid(self, canvas_shape, n_levels: int, *, src=None, src_mem=SB_MEM_HOST, src_row_pitch=0, dtype=None,
                out=None, out_mem=SB_MEM_HOST, lane=-1):
        """Nearest-neighbour x2 levels ``1 .. n_levels - 1`` of a ``(..., H, W)`` canvas (``sb_pyramid``).

        ``src=None`` takes the canvas the lane's last row-major ``fuse_region`` left on the device (no upload).
        Returns the list of level arrays, each shaped ``canvas_shape[:-2] + (h_l, w_l)`` -- views into ``out``
        (allocated here when ``None``; host output only)."""
        shapes = self.pyramid_shapes(canvas_shape, n_levels)
        if not shapes:
            return []
        planes, h, w = shapes[0][0], int(canvas_shape[-2]), int(canvas_shape[-1])
        if dtype is None:
            dtype = _pixel_dtype(src if src is not None else out)
        np_dt = np.uint8 if dtype == SB_U8 else np.uint16
        total = sum(p * a * b for p, a, b in shapes)
        if out is None:
            if out_mem != SB_MEM_HOST:
                raise ValueError("device output needs an explicit `out` address")
            out = np.empty(total, dtype=np_dt)
        self._check(self.lib.sb_pyramid(self.handle, _ptr(src) if src is not None else None, src_mem, planes, h, w,
                                        int(src_row_pitch), int(dtype), int(n_levels), _ptr(out), out_mem, lane),
                    "sb_pyramid")
        if out_mem != SB_MEM_HOST or not isinstance(out, np.ndarray):
            return out
        levels, off = [], 0
        flat = out.reshape(-1)
        for p, a, b in shapes:
            levels.append(flat[off:off + p * a * b].reshape(tuple(canvas_shape[:-2]) + (a, b)))
            off += p * a * b
        return levels


class PendingRegistration:
    """Results of ``register_pairs_async``: the ctypes buffers stay alive here until they are read."""

    def __init__(self, ctx, lane, arr, res, job, keep):
        self.ctx, self.lane, self._arr, self._res, self._job, self._keep = ctx, lane, arr, res, job, keep
        self._done = res is None

    def get(self):
        """Waits for the lane (``sb_sync``) if that has not happened yet and returns the list of result dicts."""
        if self._res is None:
            return []
        if not self._done:
            self.ctx.sync(self.lane)
            self._done = True
        return Context._pair_dicts(self._res)

    def mark_synced(self):
        self._done = True


def canvas_pitch(width: int) -> int:
    return int(load_library().sb_canvas_pitch(int(width)))
