"""CPU tests: the oracle against (a) the golden vectors produced by the unmodified
reference under shims, (b) scikit-image's known-answer tests for phase_cross_correlation."""
import hashlib

import numpy as np
import pytest
from scipy import ndimage
import scipy.fft as sfft

from conftest import RGB_GOLDENS, SMALL_GOLDENS, GOLDEN_DIR, load_golden
from oracle import pcc_ref, stitch_ref as sr, synth


def _world(shape, seed):
    return synth.make_world(shape[0], shape[1], np.random.default_rng(seed)).astype(np.float64)


def test_pcc_integer_shift_known_answer():
    # skimage test_correlation: apply (-7, 12) with fourier_shift -> expect (7, -12)
    ref = _world((128, 160), 1)
    mov = sfft.ifftn(ndimage.fourier_shift(sfft.fftn(ref), (-7, 12)))
    shift, err, _ = pcc_ref.phase_cross_correlation(ref, mov.real)
    assert tuple(shift) == (7.0, -12.0)


@pytest.mark.parametrize("shape,applied,uf", [((128, 160), (-2.4, 1.32), 100), ((1024, 214), (3, 9.2), 10),
                                              ((214, 1024), (-4.5, 2.5), 10), ((300, 157), (10.3, -7.7), 10)])
def test_pcc_subpixel_known_answer(shape, applied, uf):
    ref = _world(shape, 2)
    mov = sfft.ifftn(ndimage.fourier_shift(sfft.fftn(ref), applied)).real
    shift, _, _, det = pcc_ref.phase_cross_correlation(ref, mov, upsample_factor=uf, return_details=True)
    np.testing.assert_allclose(shift, -np.array(applied), atol=0.05 if uf == 100 else 0.1001)
    # the host-side rebuild from integer indices is bit-identical to the float path
    rebuilt = pcc_ref.shift_from_indices(det["coarse"], det["fine"], shape, uf)
    assert rebuilt.tobytes() == np.asarray(shift, dtype=np.float64).tobytes()


@pytest.mark.parametrize("n,applied", [(107, 53), (107, 54), (1024, 512), (1024, 513), (1024, 511)])
def test_pcc_wrap_rule(n, applied):
    # peak index idx = (-applied) mod n; indices beyond fix(n/2) wrap negative, exactly n/2 stays positive
    ref = _world((8, n), 3)
    mov = np.roll(ref, applied, axis=1)
    shift, _, _ = pcc_ref.phase_cross_correlation(ref, mov)
    idx = (-applied) % n
    assert shift[1] == (idx if idx <= n // 2 else idx - n)


def test_pcc_length_one_axis_is_zero():
    ref = _world((1, 64), 4)
    shift, _, _ = pcc_ref.phase_cross_correlation(ref, np.roll(ref, 3, axis=1), upsample_factor=10)
    assert shift[0] == 0 and shift[1] == -3


def test_shift_calls_match_reference():
    g = np.load(f"{GOLDEN_DIR}/shift_calls.npz")
    for i in range(int(g["n"])):
        ov = int(g[f"ov_{i}"])
        assert sr.calculate_horizontal_shift(g[f"a_{i}"], g[f"bh_{i}"], ov) == tuple(g[f"h_{i}"])
        assert sr.calculate_vertical_shift(g[f"a_{i}"], g[f"bv_{i}"], ov) == tuple(g[f"v_{i}"])


@pytest.mark.parametrize("name", SMALL_GOLDENS + RGB_GOLDENS)
def test_oracle_reproduces_reference_golden(name):
    g, st, tiles, kw = load_golden(name)
    if st.use_registration:
        sr.calculate_shifts(st, tiles)
        assert tuple(st.h_shift) == tuple(g["h_shift"])
        assert tuple(st.v_shift) == tuple(g["v_shift"])
        if st.scan_pattern == "S-Pattern":
            assert tuple(st.h_shift_rev) == tuple(g["h_shift_rev"])
            assert int(st.h_shift_rev_odd) == int(g["h_shift_rev_odd"])
    canvas = sr.stitch_region(st, tiles)
    assert canvas.shape == tuple(g["canvas_shape"])
    assert np.array_equal(canvas, g["canvas"])


def test_oracle_full_size_config0():
    """BASELINE.json configs[0]: 2x2 of 2048^2 uint16, registration + stitch (runs on CPU)."""
    g, _, _, kw = load_golden("full_2x2_2048")
    st, tiles, truth = synth.make_region(**kw)
    if hashlib.sha256(np.stack([t.pixels for t in tiles]).tobytes()).hexdigest() != str(g["input_sha"]):
        pytest.skip("synthetic generator produced different inputs than when the golden was made")
    sr.calculate_shifts(st, tiles)
    assert tuple(st.h_shift) == tuple(g["h_shift"]) == truth["h_shift"]
    assert tuple(st.v_shift) == tuple(g["v_shift"]) == truth["v_shift"]
    canvas = sr.stitch_region(st, tiles)
    assert canvas.shape == tuple(g["canvas_shape"])
    assert hashlib.sha256(canvas.tobytes()).hexdigest() == str(g["canvas_sha"])


def test_normalize_truncates_and_handles_flat():
    img = np.array([[10, 11], [12, 13]], dtype=np.uint16)
    out = sr.normalize_image(img)
    assert out.tolist() == [[0, 21845], [43690, 65535]]
    assert sr.normalize_image(np.full((4, 4), 7, np.uint16)).sum() == 0


def test_seam_crops_floor_division():
    st = sr.RegionState(tile_h=100, tile_w=100, pixel_size_um=0.5, use_registration=True,
                        h_shift=(3, -21), v_shift=(-21, -5))
    # -(-21)//2 = 10, abs(3)//2 = 1 -> 9 ; (-(-21))//2 - abs(-5)//2 = 10 - 2 = 8
    assert sr.seam_crops(st, 1, 1, 3, 3) == (9, 9, 8, 8)
    assert sr.seam_crops(st, 0, 0, 3, 3) == (0, 9, 0, 8)
    assert sr.seam_crops(st, 2, 2, 3, 3) == (9, 0, 8, 0)
