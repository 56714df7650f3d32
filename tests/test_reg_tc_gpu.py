"""Stage-by-stage checks of the tensor-core registration path (reg_tc.cu) against NumPy, through the ``sb_debug_read``
test hook: the short-axis half spectra made by ``tcgen05.mma`` (3 x tf32), the warp-level column FFT + cross-power, the
conjugate mirror columns the radix kernels downstream read.  End-to-end parity of the same path against the complex128
oracle is covered by test_register_gpu.py / test_subpixel_gpu.py (every 2048^2-tile case runs through it)."""
import os

import numpy as np
import pytest
import scipy.fft as sfft

from oracle import stitch_ref as sr
from oracle import synth

pytestmark = pytest.mark.gpu
H_DIR, V_DIR = 0, 1
F32 = 0


@pytest.fixture(scope="module")
def ctx():
    from image_stitcher_b200 import _ffi
    c = _ffi.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="module")
def tiles():
    st, t, truth = synth.make_region(rows=2, cols=2, tile_h=2048, tile_w=2048, seed=7, jitter=3)
    grid = {t_.fov: t_.pixels for t_ in t}
    return grid, truth


def frames(a, b, ov, direction):
    """The two normalised strips in the kernels' frame: long axis (1024) first, short axis (ov) second, scaled 2^-16."""
    na, nb = sr.normalize_image(a).astype(np.float64), sr.normalize_image(b).astype(np.float64)
    if direction == H_DIR:
        m = int(a.shape[0] * 0.25)
        fa, fb = na[m:-m, -ov:], nb[m:-m, :ov]
    else:
        m = int(a.shape[1] * 0.25)
        fa, fb = na[-ov:, m:-m].T, nb[:ov, m:-m].T
    return fa / 65536.0, fb / 65536.0


@pytest.mark.parametrize("direction,ov", [(H_DIR, 214), (V_DIR, 214), (H_DIR, 107), (V_DIR, 150)])
def test_tensor_core_stages_match_numpy(ctx, tiles, direction, ov):
    grid, _ = tiles
    a, b = (grid[0], grid[1]) if direction == H_DIR else (grid[0], grid[2])
    res = ctx.register_pairs([(a, b, direction)], (2048, 2048), ov, ov, precision=F32)[0]
    n, Sh, nb = ov, 1024, ov // 2 + 1
    zh = ctx.debug_read(0, 0, 2 * nb * Sh)
    assert zh.size == 2 * nb * Sh, "this strip shape did not take the tensor-core path"
    zh = zh.reshape(2, nb, Sh)
    fa, fb = frames(a, b, ov, direction)
    # ---- T1: half spectra along the short axis, stored [k][y]
    for img, f in ((0, fa), (1, fb)):
        exp = sfft.rfft(f, axis=1).T
        err = np.abs(zh[img] - exp).max() / np.abs(exp).max()
        assert err < 2e-6, (img, err)
    # ---- T2: column FFTs, cross-power (against float64 from the device's own half spectra), inverse FFT
    # (half arrays of n/2 + 1 lines when every consumer is a tensor-core stage; full arrays with the mirrored lines for
    # the radix kernels otherwise: SB_REG_NO_TC_INV / SB_REG_NO_TC_UPDFT)
    Y = ctx.debug_read(0, 1, n * Sh).reshape(-1, Sh)
    R = ctx.debug_read(0, 2, n * Sh).reshape(-1, Sh)
    assert R.shape[0] in (nb, n) and Y.shape[0] == R.shape[0]
    half = R.shape[0] == nb
    A = sfft.fft(zh[0].astype(np.complex128), axis=1)
    B = sfft.fft(zh[1].astype(np.complex128), axis=1)
    P = A * np.conj(B)
    mag = np.abs(P)
    Rn = P / np.maximum(mag, 1e-300)
    # float32 phase error grows as |P| shrinks (the spectra carry an ABSOLUTE error ~1e-7 of their largest bins): weigh
    # the error of a bin by its magnitude relative to the median bin
    weight = np.minimum(mag / np.median(mag), 1.0)
    assert (np.abs(R[:nb] - Rn) * weight).max() < 1e-3
    assert np.abs(np.abs(R[:nb]) - 1.0).max() < 1e-5
    Yn = sfft.ifft(R[:nb].astype(np.complex128), axis=1) * Sh
    assert np.abs(Y[:nb] - Yn).max() / np.abs(Yn).max() < 2e-6
    # ---- mirror columns are exact conjugates: R[n-kx][-ky] = conj(R[kx][ky]), Y[n-kx][y] = conj(Y[kx][y])
    if half:
        R = np.concatenate([R, np.zeros((n - nb, Sh), R.dtype)])
        for kx in range(1, (n - 1) // 2 + 1):
            R[n - kx] = np.conj(R[kx][(-np.arange(Sh)) % Sh])
    else:
        for kx in range(1, (n - 1) // 2 + 1):
            assert np.array_equal(R[n - kx], np.conj(R[kx][(-np.arange(Sh)) % Sh]))
            if os.environ.get("SB_REG_NO_TC_INV"):
                assert np.array_equal(Y[n - kx], np.conj(Y[kx]))
    # ---- T4: first stage of the upsampled DFT around the coarse peak, T[u][y] = sum_x conj(R[y][x]) Ex[u][x]
    uf, rs = 10, 15
    T = ctx.debug_read(0, 3, rs * Sh).reshape(rs, Sh)
    cy, cx = res["coarse"]
    if direction == V_DIR:                                       # the kernels work in the transposed frame
        cy, cx = cx, cy
    cxw = cx - n if cx > n // 2 else cx
    off = rs // 2 - cxw * uf
    Ex = np.exp(-2j * np.pi * (np.arange(rs) - off)[:, None] * sfft.fftfreq(n, uf)[None, :])
    Tn = Ex @ np.conj(R.astype(np.complex128))                   # R is stored [x][y]
    assert np.abs(T - Tn).max() / np.abs(Tn).max() < 5e-6
    # ---- and the chain built on them lands where the complex128 oracle does (the narrower strips do not reach the
    # 205-pixel overlap of these tiles: their correlation is noise, an argmax float32 need not reproduce)
    if ov == 214:
        fn = sr.calculate_horizontal_shift if direction == H_DIR else sr.calculate_vertical_shift
        ints, shift, det = fn(a, b, ov, upsample_factor=10, return_details=True)
        assert res["coarse"] == det["coarse"] and res["fine"] == det["fine"] and (res["dy"], res["dx"]) == ints


def test_tensor_core_path_batch_of_wells_matches_truth(ctx):
    """A batch large enough for several sub-batches and streams: every pair of 6 wells recovers the known drift."""
    import torch
    from image_stitcher_b200 import _ffi
    from image_stitcher_b200.plate import PlateSpec, make_plate, well_pairs
    spec = PlateSpec(wells=6, rows=3, cols=3, tile_h=2048, tile_w=2048, channels=1, reg_channel=0, jitter=3, seed=19)
    plate = make_plate(spec, device="cuda:0", with_flat=False)
    pairs, kinds = [], []
    for w in range(spec.wells):
        p, k = well_pairs(spec, lambda r, c, ch, z, w=w: plate.pool[w, r, c, ch, z].data_ptr())
        pairs += p
        kinds += [(w, kk) for kk in k]
    torch.cuda.synchronize()
    ovx, ovy = spec.strip_overlaps()
    res = ctx.register_pairs(pairs, (2048, 2048), ovx, ovy, mem=_ffi.SB_MEM_DEVICE)
    assert all((r["dy"], r["dx"]) == plate.truth[w][k] for r, (w, k) in zip(res, kinds))
    assert all(r["fine"] == (7, 7) for r in res)
