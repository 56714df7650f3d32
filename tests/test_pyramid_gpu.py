"""GPU parity of the multiscale levels (``sb_pyramid``): the reference saves a region through ome_zarr's
``Scaler(method="nearest")`` (stitcher_process.py:1061-1062), i.e. ``level[..., ::2, ::2]`` per level.
Byte movement only, so the bar is bit-exact against NumPy slicing."""
import multiprocessing as mp
import os

import numpy as np
import pytest

from conftest import load_golden
from image_stitcher_b200 import ome_zarr_writer as ozw
from image_stitcher_b200.stitcher_parameters import StitchingParameters
from oracle import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from image_stitcher_b200 import _ffi
    c = _ffi.Context(0)
    yield c
    c.close()


def _slices(canvas, n_levels):
    out, level = [], canvas
    for _ in range(1, n_levels):
        level = level[..., ::2, ::2]
        out.append(level)
    return out


@pytest.mark.parametrize("dtype", [np.uint16, np.uint8])
@pytest.mark.parametrize("shape", [(1, 2, 1, 301, 517), (1, 1, 1, 64, 64), (3, 1023, 7), (1, 1, 9), (2, 2, 2)])
def test_levels_from_a_host_canvas_match_numpy_slicing(ctx, dtype, shape):
    rng = np.random.default_rng(hash((shape, np.dtype(dtype).itemsize)) & 0xFFFF)
    canvas = rng.integers(0, np.iinfo(dtype).max + 1, size=shape, dtype=dtype)
    n_levels = 6
    got = ctx.pyramid(canvas.shape, n_levels, src=canvas)
    want = _slices(canvas, n_levels)
    assert len(got) == len(want) == n_levels - 1
    for a, b in zip(got, want):
        assert a.shape == b.shape and a.dtype == b.dtype
        assert np.array_equal(a, b)


def test_source_row_pitch_and_one_level_is_a_no_op(ctx):
    rng = np.random.default_rng(5)
    padded = rng.integers(0, 65536, size=(2, 100, 256), dtype=np.uint16)      # rows 256 apart, 199 used
    got = ctx.pyramid((2, 100, 199), 3, src=padded, src_row_pitch=256)
    want = _slices(padded[:, :, :199], 3)
    assert all(np.array_equal(a, b) for a, b in zip(got, want))
    assert ctx.pyramid((2, 100, 199), 1, src=padded, src_row_pitch=256) == []


@pytest.mark.parametrize("dtype", [np.uint16, np.uint8])
def test_levels_from_the_resident_canvas_of_the_last_fuse(ctx, dtype):
    """src=None: the levels come from the canvas sb_fuse_region left on the device (padded pitch, no upload)."""
    from image_stitcher_b200 import _ffi
    rng = np.random.default_rng(11)
    H = W = 128
    tiles = [rng.integers(0, np.iinfo(dtype).max + 1, size=(H, W), dtype=dtype) for _ in range(4)]
    job = [(tiles[0], 0, 0, 0, 0, 0, 0, 0, 0), (tiles[1], 115, 2, 0, 0, 0, 0, 0, 0),
           (tiles[2], 3, 117, 1, 0, 0, 0, 0, 0), (tiles[3], 118, 119, 1, 0, 0, 0, 0, 0)]
    shape = (1, 2, 1, 249, 247)                       # width not a multiple of the device pitch
    out = np.empty(shape, dtype=dtype)
    ctx.fuse_region(job, (H, W), shape[1:], out=out)
    before = ctx.kernel_launches
    got = ctx.pyramid(shape, 4, dtype=_ffi._pixel_dtype(out))
    assert ctx.kernel_launches == before + 3         # one launch per level, nothing else
    for a, b in zip(got, _slices(out, 4)):
        assert np.array_equal(a, b)


def test_errors(ctx):
    from image_stitcher_b200 import _ffi
    c = _ffi.Context(0)
    try:
        with pytest.raises(RuntimeError, match="no resident canvas"):
            c.pyramid((1, 8, 8), 2, dtype=_ffi.SB_U16)
        tile = np.zeros((64, 64), np.uint16)
        out = np.empty((1, 1, 1, 64, 64), np.uint16)
        c.fuse_region([(tile, 0, 0, 0, 0, 0, 0, 0, 0)], (64, 64), (1, 1, 64, 64), out=out)
        with pytest.raises(RuntimeError, match="resident canvas is"):
            c.pyramid((1, 1, 1, 64, 32), 2, dtype=_ffi.SB_U16)
        c.normalize(tile)                                   # stages through lane 0's canvas buffer: the resident canvas is gone
        with pytest.raises(RuntimeError, match="no resident canvas"):
            c.pyramid((1, 1, 1, 64, 64), 2, dtype=_ffi.SB_U16)
        with pytest.raises(RuntimeError, match="bad canvas shape"):
            c.pyramid((1, 0, 8), 2, src=np.zeros(8, np.uint16))
    finally:
        c.close()
    assert _ffi.load_library().sb_pyramid_elems(2, 5, 7, 3) == 2 * (3 * 4 + 2 * 2)
    assert _ffi.load_library().sb_pyramid_elems(2, 0, 7, 3) == -1


def test_stitcher_process_writes_gpu_made_levels(tmp_path, monkeypatch):
    """``run()`` on the plain path (row-major canvas; SB_NO_FAST_IO=1 -- the chunk-ordered fast path is covered by
    test_run_fast_path_equals_the_plain_path_level_by_level) hands the GPU-made levels to the OME-Zarr writer; every
    stored level equals the slicing of level 0."""
    from image_stitcher_b200 import geometry as geo
    monkeypatch.setenv("SB_NO_FAST_IO", "1")
    from image_stitcher_b200.stitcher_process import StitcherProcess
    g, st, tiles, kw = load_golden("reg_2x2_mono")
    root = str(tmp_path / "acq")
    synth.write_squid_layout(root, {"A1": tiles})
    monkeypatch.setattr(geo, "pyramid_levels", lambda *a, **k: 4)
    seen = {}
    real = ozw.write_ome_zarr

    def spy(path, data, **kwargs):
        seen["levels"] = kwargs.get("levels")
        return real(path, data, **kwargs)
    monkeypatch.setattr(ozw, "write_ome_zarr", spy)
    p = StitchingParameters(input_folder=root, use_registration=st.use_registration, apply_flatfield=st.apply_flatfield,
                            scan_pattern=st.scan_pattern, registration_channel=st.registration_channel)
    s = StitcherProcess(p, mp.Queue(), mp.Queue(), mp.Queue(), mp.Event())
    s.run()
    kind, (path, _) = s.complete_queue.get(timeout=5)
    assert kind == "complete" and os.path.isdir(path)
    assert seen["levels"] is not None and len(seen["levels"]) == 3
    level0 = ozw.read_ome_zarr_level(path, 0)
    assert np.array_equal(level0, g["canvas"])
    for l, want in enumerate(_slices(level0, 4), start=1):
        assert np.array_equal(ozw.read_ome_zarr_level(path, l), want), l
