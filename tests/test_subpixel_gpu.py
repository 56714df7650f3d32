"""Sub-pixel / half-pixel parity of the CUDA registration chain against the complex128 oracle.

The reference returns ``round(shift[1] - Sw)`` (Python half-even round of a value on a 1/upsample grid,
stitcher_process.py:683-685, 706-708), so one step of the 15x15 fine argmax at x.45 / x.5 / x.55, or a coarse
argmax that lands on the other pixel of a half-pixel tie, changes the INTEGER shift.  These tests shift the
moving tile by fractions of a pixel (``scipy.ndimage.fourier_shift``) at the strip shapes of the BASELINE
configurations and demand the same coarse index, fine index, float64 shift (bit-equal) and rounded integers as
``oracle/pcc_ref.py`` -- unconditionally in SB_PREC_F64 and SB_PREC_AUTO, and in SB_PREC_F32 whenever the result's
own margins say the argmax was not a near-tie (that is the contract AUTO's float64 redo is built on)."""
import numpy as np
import pytest
import scipy.fft as sfft
from scipy import ndimage

from oracle import stitch_ref as sr
from oracle import synth

pytestmark = pytest.mark.gpu
H_DIR, V_DIR = 0, 1
F32, F64, AUTO = 0, 1, 2
TIE = 1e-4        # include/stitchb200.h: relative margin below which SB_PREC_AUTO repeats a pair in float64

FRACTIONS = [(0.04, 0.45), (0.05, 0.5), (0.45, 0.55), (0.5, 0.95), (0.55, 0.04), (0.95, 0.05), (0.5, 0.5), (0.25, 0.75)]


@pytest.fixture(scope="module")
def ctx():
    from image_stitcher_b200 import _ffi
    c = _ffi.Context(0)
    yield c
    c.close()


def _shifted_window(world, y0, x0, h, w, fy, fx, pad=32):
    """world[y0:y0+h, x0:x0+w] with its content displaced by (fy, fx) pixels (Fourier shift of a padded window)."""
    win = world[y0 - pad:y0 + h + pad, x0 - pad:x0 + w + pad]
    sh = sfft.ifft2(ndimage.fourier_shift(sfft.fft2(win), (fy, fx))).real
    return sh[pad:pad + h, pad:pad + w]


def fractional_pairs(H, W, tov, ov, fractions, seed, noise=25.0):
    """Reference tile +, per (fy, fx), a right and a bottom neighbour whose STRIP content (the only pixels the phase
    correlation reads; the rest of the tile only feeds the whole-tile min/max) is displaced by a fraction of a pixel
    plus an integer jitter from the nominal ``tov``-pixel overlap.  Returns [(a, b, dir), ...]."""
    rng = np.random.default_rng(seed)
    pad = 48
    world = synth.make_world(2 * H + 2 * pad, 2 * W + 2 * pad, rng).astype(np.float64)
    a = np.clip(world[pad:pad + H, pad:pad + W] + rng.normal(0, noise, (H, W)), 0, 65535).astype(np.uint16)
    my, mx = int(H * 0.25), int(W * 0.25)
    out = []
    for k, (fy, fx) in enumerate(fractions):
        jy, jx = (k % 3) - 1, ((k // 3) % 3) - 1
        y0, x0 = pad + jy, pad + W - tov + jx                                    # right neighbour
        bh = world[y0:y0 + H, x0:x0 + W].copy()
        bh[my:H - my, :ov] = _shifted_window(world, y0 + my, x0, H - 2 * my, ov, fy, fx)
        y0, x0 = pad + H - tov + jy, pad + jx                                    # bottom neighbour
        bv = world[y0:y0 + H, x0:x0 + W].copy()
        bv[:ov, mx:W - mx] = _shifted_window(world, y0, x0 + mx, ov, W - 2 * mx, fy, fx)
        out.append((a, np.clip(bh + rng.normal(0, noise, (H, W)), 0, 65535).astype(np.uint16), H_DIR))
        out.append((a, np.clip(bv + rng.normal(0, noise, (H, W)), 0, 65535).astype(np.uint16), V_DIR))
    return out


def oracle(a, b, ov, d, uf):
    fn = sr.calculate_horizontal_shift if d == H_DIR else sr.calculate_vertical_shift
    return fn(a, b, ov, upsample_factor=uf, return_details=True)


def clear_margins(r, uf):
    return (r["peak"] - r["second"] > TIE * r["peak"]) and (uf == 1 or r["fine_peak"] - r["fine_second"] > TIE * r["fine_peak"])


def check(ctx, job, tile_shape, ov, uf, stats):
    exp = [oracle(a, b, ov, d, uf) for a, b, d in job]
    for prec in (F64, AUTO, F32):
        res = ctx.register_pairs(job, tile_shape, ov, ov, upsample_factor=uf, precision=prec)
        for r, (ints, shift, det), (_, _, d) in zip(res, exp, job):
            strict = prec != F32 or clear_margins(r, uf)
            if strict:
                assert r["coarse"] == det["coarse"], (prec, d, r, det)
                assert r["fine"] == (det["fine"] if uf > 1 else (-1, -1)), (prec, d, r, det)
                assert np.array(r["shift"]).tobytes() == np.asarray(shift, np.float64).tobytes()
                assert (r["dy"], r["dx"]) == ints
            else:
                stats["f32_near_ties"] += 1
                assert np.abs(np.array(r["shift"]) - shift).max() <= 1.0 / uf + 1e-12      # north_star tolerance
            if prec == AUTO:
                stats["auto_pairs"] += 1
                stats["auto_redone"] += int(r["precision"] == F64)
                if r["precision"] == F32:
                    assert clear_margins(r, uf)          # whatever stayed in float32 had clear margins
            # second-largest values are ordered and the band runner-up cannot exceed the second
            assert r["peak"] >= r["second"] >= 0 and r["second"] >= r["runner_up"] - 1e-6
            if uf > 1:
                assert r["fine_peak"] >= r["fine_second"] >= 0


@pytest.mark.parametrize("uf", [10, 100])
def test_fractional_shifts_2048_tiles(ctx, uf):
    """Strips 1024x214 and 214x1024 (BASELINE configs[0..3])."""
    stats = dict(f32_near_ties=0, auto_pairs=0, auto_redone=0)
    job = fractional_pairs(2048, 2048, 205, 214, FRACTIONS if uf == 10 else FRACTIONS[:4], seed=71 + uf)
    check(ctx, job, (2048, 2048), 214, uf, stats)
    print("fractional 2048:", stats)


def test_fractional_shifts_3000_tiles(ctx):
    """Strips 1500x314 and 314x1500 (BASELINE configs[4]: 314 = 2 * 157)."""
    stats = dict(f32_near_ties=0, auto_pairs=0, auto_redone=0)
    job = fractional_pairs(3000, 3000, 300, 314, [(0.45, 0.5), (0.5, 0.05), (0.55, 0.95)], seed=73)
    check(ctx, job, (3000, 3000), 314, 10, stats)
    print("fractional 3000:", stats)


@pytest.mark.parametrize("uf", [10, 4, 1])
def test_fractional_shifts_small_odd_strips(ctx, uf):
    """Odd strip extents (no Nyquist bin): 150 x 45 and 45 x 150."""
    stats = dict(f32_near_ties=0, auto_pairs=0, auto_redone=0)
    job = fractional_pairs(300, 300, 40, 45, FRACTIONS, seed=79)
    check(ctx, job, (300, 300), 45, uf, stats)
    print("fractional small:", stats)


def _tie_tiles(H, W, ov, bump, direction, seed, scale_a=20000):
    """Two tiles whose strips are scale_a * s and K * (s + roll(s, 1)) with one pixel raised by ``bump``: in exact
    arithmetic the cross-power phase is that of a half-pixel shift (two EQUAL peaks on neighbouring pixels for an odd
    extent); the bump separates them by ~1e-7 .. 1e-5 of the peak -- far above complex128 rounding, at or below float32's.
    Both tiles hold a 0 and a 65535 outside the strips, so normalize_image is the identity on them."""
    rng = np.random.default_rng(seed)
    m = int((H if direction == H_DIR else W) * 0.25)
    n_long = (H if direction == H_DIR else W) - 2 * m
    s = rng.integers(0, 2, (n_long, ov)).astype(np.int64)
    t = 30000 * (s + np.roll(s, 1, axis=1))
    t[5, 7] += bump
    s = s * scale_a
    A = np.zeros((H, W), np.uint16)
    B = np.zeros((H, W), np.uint16)
    if direction == H_DIR:
        A[m:H - m, W - ov:] = s
        B[m:H - m, :ov] = t
        A[0, 1] = B[0, 1] = 65535
    else:
        A[H - ov:, m:W - m] = s.T
        B[:ov, m:W - m] = t.T
        A[0, 1] = B[H - 1, 1] = 65535
    return A, B


@pytest.mark.parametrize("bump", [1, 40, 300])
def test_auto_repeats_coarse_near_ties_in_float64(ctx, bump):
    H = W = 256
    ov = 33
    job = []
    for d in (H_DIR, V_DIR):
        A, B = _tie_tiles(H, W, ov, bump, d, seed=5 + bump)
        job.append((A, B, d))
    exp = [oracle(a, b, ov, d, 10) for a, b, d in job]
    for prec in (AUTO, F64):
        res = ctx.register_pairs(job, (H, W), ov, ov, precision=prec)
        for r, (ints, shift, det) in zip(res, exp):
            assert r["precision"] == F64                              # AUTO saw the thin margin and repeated the pair
            assert 0 <= r["peak"] - r["second"] <= TIE * r["peak"]
            assert r["coarse"] == det["coarse"] and r["fine"] == det["fine"]
            assert (r["dy"], r["dx"]) == ints
            assert np.array(r["shift"]).tobytes() == np.asarray(shift, np.float64).tobytes()
    # plain float32 reports the same thin margin (so a caller can see it) even if it picks the other pixel
    for r in ctx.register_pairs(job, (H, W), ov, ov, precision=F32):
        assert r["precision"] == F32 and r["peak"] - r["second"] <= TIE * r["peak"]


def test_auto_repeats_pairs_with_strips_of_very_different_magnitude(ctx):
    """The radix path transforms both strips as ONE complex array (a + i b); a strip 30000 x fainter than its partner is
    then recovered by cancellation and float32 keeps ~3 digits of its spectrum -- the correlation is off by ~2e-4 of the
    peak, more than the near-tie threshold can see.  AUTO detects the magnitude ratio itself and repeats the pair."""
    H = W = 256
    ov = 33
    job = [_tie_tiles(H, W, ov, 300, d, seed=11, scale_a=1) + (d,) for d in (H_DIR, V_DIR)]
    exp = [oracle(a, b, ov, d, 10) for a, b, d in job]
    for r, (ints, shift, det) in zip(ctx.register_pairs(job, (H, W), ov, ov, precision=AUTO), exp):
        assert r["precision"] == F64
        assert r["coarse"] == det["coarse"] and r["fine"] == det["fine"] and (r["dy"], r["dx"]) == ints


def test_runner_up_excludes_the_peak_band(ctx):
    """A pair with two correlation peaks (the moving strip is the sum of two displaced copies): `second` is the
    neighbouring lobe or the other peak, `runner_up` is the largest value at least two lines away from the peak
    along the strip's long axis -- here the second copy, 9 lines away."""
    rng = np.random.default_rng(17)
    H, W, ov = 256, 320, 40
    world = synth.make_world(H + 64, 2 * W + 64, rng)
    a = np.clip(world[20:20 + H, 20:20 + W], 0, 30000).astype(np.uint16)
    x0 = 20 + W - ov + 2
    b1 = world[20:20 + H, x0:x0 + W]
    b2 = world[29:29 + H, x0:x0 + W]
    b = np.clip(0.6 * b1 + 0.4 * b2, 0, 65535).astype(np.uint16)
    r = ctx.register_pairs([(a, b, H_DIR)], (H, W), ov, ov, precision=F64)[0]
    ints, shift, det = oracle(a, b, ov, H_DIR, 10)
    assert r["coarse"] == det["coarse"] and (r["dy"], r["dx"]) == ints
    P = sfft.fftn(sr.normalize_image(a)[64:-64, -ov:]) * np.conj(sfft.fftn(sr.normalize_image(b)[64:-64, :ov]))
    P /= np.maximum(np.abs(P), 100 * np.finfo(np.float64).eps)
    cc = np.abs(sfft.ifftn(P))
    cy = det["coarse"][0]
    band = [(cy - 1) % cc.shape[0], cy, (cy + 1) % cc.shape[0]]
    outside = np.delete(cc, band, axis=0).max()
    flat = np.sort(cc.ravel())
    assert abs(r["runner_up"] - outside) <= 1e-6 and abs(r["second"] - flat[-2]) <= 1e-6 and abs(r["peak"] - flat[-1]) <= 1e-6
