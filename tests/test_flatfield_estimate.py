"""Flat-field ESTIMATOR extension (``sb_estimate_flatfield``).  The reference fits
BaSiC (third-party, not available offline), so this row has no reference parity: the CUDA path is compared with the
repo's own definition (oracle/flatfield_ref.py), and the definition with the vignette the tiles were made with."""
import multiprocessing as mp

import numpy as np
import pytest

from oracle import flatfield_ref as fr


def _sample(n=20, h=256, w=320, dtype=np.uint16, seed=0):
    """Smooth background + sparse bright blobs, multiplied by a radial vignette of mean 1."""
    from scipy.ndimage import gaussian_filter
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    r2 = ((yy - h / 2) ** 2 + (xx - w / 2) ** 2) / (h * h / 4 + w * w / 4)
    vig = 1.1 - 0.4 * r2
    vig /= vig.mean()
    top = float(np.iinfo(dtype).max)
    tiles = []
    for _ in range(n):
        base = (0.01 + 0.005 * gaussian_filter(rng.random((h, w)), 8)) * top
        blobs = gaussian_filter((rng.random((h, w)) > 0.9995).astype(float), 3) * 0.6 * top
        img = (base + blobs) * vig + rng.normal(0, 0.0002 * top, (h, w))
        tiles.append(np.clip(img, 0, top).astype(dtype))
    return np.array(tiles), vig


def test_oracle_recovers_the_vignette_and_handles_edge_cases():
    tiles, vig = _sample()
    f = fr.estimate_flatfield(tiles, grid=64)
    assert f.dtype == np.float32 and f.shape == vig.shape
    assert abs(float(f.mean()) - 1.0) < 1e-3
    err = np.abs(f / vig - 1.0)
    assert err.mean() < 0.01 and err.max() < 0.08          # worst at the corners (smoothing + clamped interpolation)
    # the median ignores what a mean would not: one saturated tile changes nothing much
    tiles2 = tiles.copy()
    tiles2[0] = 65535
    assert np.abs(fr.estimate_flatfield(tiles2, grid=64) / f - 1.0).max() < 0.02
    assert np.array_equal(fr.estimate_flatfield(np.zeros((3, 16, 16), np.uint16)), np.ones((16, 16), np.float32))
    assert fr.estimate_flatfield(tiles[:1, :37, :53], grid=128).shape == (37, 53)      # grid clamped to the tile
    assert list(fr.cell_edges(10, 4)) == [0, 3, 5, 8, 10]
    w = fr.gaussian_weights(2.0)
    assert len(w) == 13 and abs(w.sum() - 1.0) < 1e-15 and np.allclose(w, w[::-1])
    a = np.arange(5, dtype=np.float64)[None, :]
    assert np.allclose(fr.smooth_axis(a, np.array([0.25, 0.5, 0.25]), 1), [[0.25, 1, 2, 3, 3.75]])   # a | a b c d e | e


@pytest.mark.gpu
@pytest.mark.parametrize("dtype,shape,grid,n", [(np.uint16, (256, 320), 32, 20), (np.uint16, (250, 301), 128, 7),
                                                (np.uint8, (128, 256), 16, 9), (np.uint8, (97, 131), 64, 4),
                                                (np.uint16, (64, 64), 1, 3), (np.uint16, (2048, 2048), 128, 6)])
def test_cuda_estimate_matches_the_oracle(dtype, shape, grid, n):
    from image_stitcher_b200 import _ffi
    tiles, _ = _sample(n=n, h=shape[0], w=shape[1], dtype=dtype, seed=grid)
    tiles[n - 1] = 0                                     # an all-black tile is dropped, not divided by
    c = _ffi.Context(0)
    try:
        before = c.kernel_launches
        got = c.estimate_flatfield(list(tiles), grid=grid, sigma=2.0)
        assert c.kernel_launches == before + 7
    finally:
        c.close()
    want = fr.estimate_flatfield(tiles, grid=grid, sigma=2.0)
    assert got.dtype == np.float32 and got.shape == want.shape
    assert np.allclose(got, want, rtol=2e-6, atol=0)     # float64 on both sides, float32 at the end


@pytest.mark.gpu
def test_cuda_estimate_edge_cases_and_errors():
    from image_stitcher_b200 import _ffi
    c = _ffi.Context(0)
    try:
        assert np.array_equal(c.estimate_flatfield([np.zeros((32, 40), np.uint16)] * 3), np.ones((32, 40), np.float32))
        tiles, _ = _sample(n=3, h=64, w=64)
        assert np.allclose(c.estimate_flatfield(list(tiles), sigma=0.0), fr.estimate_flatfield(tiles, sigma=0.0), rtol=2e-6)
        with pytest.raises(RuntimeError, match="at most 128 tiles"):
            c.estimate_flatfield([tiles[0]] * 129)
        with pytest.raises(ValueError):
            c.estimate_flatfield([tiles[0], tiles[1][:32]])
    finally:
        c.close()


@pytest.mark.gpu
def test_run_with_apply_flatfield_estimates_fields_when_basicpy_is_absent(tmp_path):
    """``-ff`` end to end without BaSiCPy: ``run()`` estimates one field per channel on the GPU, the canvas is the
    oracle's paste of the tiles divided by exactly those fields."""
    pytest.importorskip("scipy")
    try:
        import basicpy  # noqa: F401
        pytest.skip("BaSiCPy is installed: the reference's own fit is used")
    except ImportError:
        pass
    from conftest import load_golden
    from image_stitcher_b200 import ome_zarr_writer as ozw
    from image_stitcher_b200.stitcher_parameters import StitchingParameters
    from image_stitcher_b200.stitcher_process import StitcherProcess
    from oracle import stitch_ref as sr, synth
    g, st, tiles, kw = load_golden("coord_2x2_plain")
    root = str(tmp_path / "acq")
    synth.write_squid_layout(root, {"A1": tiles})
    p = StitchingParameters(input_folder=root, use_registration=False, apply_flatfield=True)
    s = StitcherProcess(p, mp.Queue(), mp.Queue(), mp.Queue(), mp.Event())
    s.run()
    kind, (path, _) = s.complete_queue.get(timeout=5)
    assert kind == "complete"
    assert sorted(s.flatfields) == list(range(st.num_c))
    per_channel = {}
    for t in tiles:
        per_channel.setdefault(t.channel, []).append(t.pixels)
    for ch, planes in per_channel.items():
        want = fr.estimate_flatfield(np.array(planes))           # the sample is every tile of the channel (order-free median)
        assert np.allclose(s.flatfields[st.monochrome_channels.index(ch)], want, rtol=2e-6)
    st.apply_flatfield = True
    st.flatfields = {c: np.asarray(f) for c, f in s.flatfields.items()}
    assert np.array_equal(ozw.read_ome_zarr_level(path, 0), sr.stitch_region(st, tiles))


@pytest.mark.gpu
def test_get_flatfields_splits_rgb_tiles_into_planes(tmp_path):
    """8-bit RGB camera tiles: one estimated field per colour plane, keyed like the reference (``<ch>_R/_G/_B``, :557-565)."""
    try:
        import basicpy  # noqa: F401
        pytest.skip("BaSiCPy is installed: the reference's own fit is used")
    except ImportError:
        pass
    from conftest import load_golden
    from image_stitcher_b200.stitcher_parameters import StitchingParameters
    from image_stitcher_b200.stitcher_process import StitcherProcess
    from oracle import synth
    g, st, tiles, kw = load_golden("coord_2x3_rgb_u8")
    root = str(tmp_path / "acq")
    synth.write_squid_layout(root, {"A1": tiles})
    s = StitcherProcess(StitchingParameters(input_folder=root, apply_flatfield=True), mp.Queue(), mp.Queue(), mp.Queue(), mp.Event())
    try:
        s.get_timepoints()
        s.extract_acquisition_parameters()
        s.get_pixel_size()
        s.parse_acquisition_metadata()
        s.get_flatfields()
        assert sorted(s.flatfields) == [0, 1, 2] and len(s.monochrome_channels) == 3
        for i, suffix in enumerate("RGB"):
            idx = [k for k, name in enumerate(s.monochrome_channels) if name.endswith("_" + suffix)][0]
            want = fr.estimate_flatfield(np.array([t.pixels[:, :, i] for t in tiles]))
            assert s.flatfields[idx].shape == want.shape and np.allclose(s.flatfields[idx], want, rtol=2e-6)
    finally:
        s.cleanup()
