"""Exhaustive on-device proofs of the two primitives the bit-exactness claims rest on.

* ``stretch_px`` -- the integer form of ``normalize_image`` (stitcher_process.py:844-855) that the registration
  kernels fuse into the strip load -- against the float64 expression, for EVERY (v - min, max - min) pair;
* ``div2_rn`` + ``trunc_sat_pack`` / ``round_sat_pack`` -- the packed float32 divide and truncation of the paste kernels
  (``apply_flatfield_correction``, :828-842) -- against IEEE ``a / b`` and ``trunc(clip(.))`` for every uint16 numerator
  and every float32 flat value of a binade, at the two ends of the accepted field range and around 1."""
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from image_stitcher_b200 import _ffi
    c = _ffi.Context(0)
    yield c
    c.close()


@pytest.mark.parametrize("maxval", [255, 65535])
def test_integer_stretch_equals_float64_for_every_pixel_and_range(ctx, maxval):
    """Every form the kernels run: ``stretch_px`` (radix path), ``stretch_bits`` and ``stretch_core`` (tensor-core converters,
    interior / edge chunks) -- each with its hand-over of exact quotients to the float64 sequence."""
    checked, bad, first, fallbacks = ctx.selftest(0, maxval)
    assert checked == maxval * (maxval + 3) // 2              # 2 147 581 950 pairs for uint16
    assert bad == 0, f"{bad} mismatches, first at b={first >> 32}, a={first & 0xffffffff}"
    assert 0 < fallbacks < checked // 10                      # exact quotients exist and are rare (4.5 % of the 8-bit pairs, far fewer of the 16-bit ones)"


@pytest.mark.parametrize("expo", list(range(-5, 20)))
def test_packed_divide_equals_ieee_for_every_numerator_and_mantissa(ctx, expo):
    """Every binade of the accepted field range [2^-5, 2^20): the quotient bits are scale-invariant, but which quotients sit
    next to an integer -- what ``trunc_sat_pack`` / ``round_sat_pack`` (denormal-scale products) must get right -- depends
    on the binade.  (A divide without the Newton step passed 23 of the 25 binades and failed ONE case in each of 2^-2 and
    2^-1: r2 call 28.)"""
    checked, bad, first, _ = ctx.selftest(1, expo)
    assert checked == (1 << 23) * 65536
    assert bad == 0, f"{bad} mismatches, first at mantissa={first >> 16}, a={first & 0xffff}"


def test_tensor_core_building_blocks(ctx):
    """tcgen05.mma.kind::tf32 with the 3-term split, operands in the K-major no-swizzle layout the registration kernels
    write, accumulators read back from TMEM: float32-grade agreement with a float64 product.  The single-term variant
    shows what the split buys (errors ~2^-11 instead of ~2^-21)."""
    checked, bad, err3, flag = ctx.selftest(2, 0)
    assert flag == 0, "tensor pipeline did not complete"
    assert checked == 128 * 112 and bad == 0, f"{bad} elements off, max error {err3 * 1e-12:.3e}"
    _, bad1, err1, flag1 = ctx.selftest(2, 2)
    assert flag1 == 0 and bad1 == 0
    print(f"3 x tf32 max error {err3 * 1e-12:.2e}; 1 x tf32 max error {err1 * 1e-12:.2e}")
    assert err3 * 50 < err1


def test_selftest_rejects_unknown_requests(ctx):
    with pytest.raises(RuntimeError):
        ctx.selftest(7, 0)
    with pytest.raises(RuntimeError):
        ctx.selftest(0, 1000)
    with pytest.raises(RuntimeError):
        ctx.selftest(1, -9)
