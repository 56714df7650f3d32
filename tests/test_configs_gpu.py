"""Parity at the shapes of the other BASELINE.json configurations (they are parity-test cases, not bench lines):

* configs[1]  single well 5x5 tiles, 3 channels, flatfield on, registration on the 488 nm channel;
* configs[3]  2x2 tiles per well, registration + flatfield + feathered blend (extension, +-1 LSB vs oracle/blend_ref);
* configs[4]  3000x3000 tiles, 5-plane z-stack, upsample_factor 10 -> strips 1500x314 / 314x1500
              (1500 = 2^2 * 3 * 5^3, 314 = 2 * 157 with 157 prime).

Tile sizes are reduced where only the grid/channels matter; the FFT-size-critical case runs at the real 3000^2."""
import numpy as np
import pytest

from oracle import stitch_ref as sr
from oracle import synth

pytestmark = pytest.mark.gpu
H_DIR, V_DIR = 0, 1


@pytest.fixture(scope="module")
def ctx():
    from image_stitcher_b200 import _ffi
    c = _ffi.Context(0)
    yield c
    c.close()


def _oracle_pair(a, b, ov, direction, uf=10):
    fn = sr.calculate_horizontal_shift if direction == H_DIR else sr.calculate_vertical_shift
    ints, shift, det = fn(a, b, ov, upsample_factor=uf, return_details=True)
    return ints, shift, det


def test_config4_strip_shapes_3000px_tiles(ctx):
    """1500x314 and 314x1500 strips: radices 157, 5, 5, 5, 4, 3, 2 -- integer shifts and peak indices bit-exact
    against the complex128 oracle, in float32, float64 and auto precision."""
    from image_stitcher_b200 import geometry as geo
    st, tiles, truth = synth.make_region(rows=2, cols=2, tile_h=3000, tile_w=3000, seed=41, jitter=3, num_z=1)
    xs = sorted(set(t.x_mm for t in tiles))
    ys = sorted(set(t.y_mm for t in tiles))
    ovx, ovy = geo.strip_overlaps(3000, 3000, xs, ys, st.pixel_size_um, st.pixel_binning)
    assert (ovx, ovy) == (314, 314)
    lut = {(t.x_mm, t.y_mm): t.pixels for t in tiles}
    job = [(lut[(xs[0], ys[0])], lut[(xs[1], ys[0])], H_DIR), (lut[(xs[0], ys[0])], lut[(xs[0], ys[1])], V_DIR),
           (lut[(xs[0], ys[1])], lut[(xs[1], ys[1])], H_DIR), (lut[(xs[1], ys[0])], lut[(xs[1], ys[1])], V_DIR)]
    exp = [_oracle_pair(a, b, 314, d) for a, b, d in job]
    for precision in (0, 1, 2):
        res = ctx.register_pairs(job, (3000, 3000), ovx, ovy, precision=precision)
        for r, (ints, shift, det) in zip(res, exp):
            assert (r["dy"], r["dx"]) == ints
            assert r["coarse"] == det["coarse"] and r["fine"] == det["fine"]
            assert np.array_equal(np.array(r["shift"]), shift)
    assert (res[0]["dy"], res[0]["dx"]) == truth["h_shift"] and (res[1]["dy"], res[1]["dx"]) == truth["v_shift"]


def test_config1_5x5_three_channels_flatfield_registration_channel(ctx):
    """5x5 grid, 3 fluorescence channels, flat-field on, registration on the 488 nm channel: the reference flow
    (centre pairs -> lattice -> registered paste) through the C ABI equals the oracle's canvas bit for bit."""
    from image_stitcher_b200 import _ffi
    from image_stitcher_b200 import geometry as geo
    chans = ("Fluorescence 405 nm Ex", "Fluorescence 488 nm Ex", "Fluorescence 561 nm Ex")
    st, tiles, truth = synth.make_region(rows=5, cols=5, tile_h=320, tile_w=384, seed=52, jitter=2, channels=chans,
                                         use_registration=True, apply_flatfield=True,
                                         registration_channel="Fluorescence 488 nm Ex")
    st = sr.calculate_shifts(st, tiles)
    exp = sr.stitch_region(st, tiles)
    xs = sorted(set(t.x_mm for t in tiles))
    ys = sorted(set(t.y_mm for t in tiles))
    ovx, ovy = geo.strip_overlaps(st.tile_w, st.tile_h, xs, ys, st.pixel_size_um, st.pixel_binning)
    lut = {(t.x_mm, t.y_mm): t.pixels for t in tiles if t.channel == "Fluorescence 488 nm Ex"}
    plan, _ = geo.center_pairs(xs, ys, False)
    res = ctx.register_pairs([(lut[a], lut[b], H_DIR if k == "h" else V_DIR) for k, a, b in plan],
                             (st.tile_h, st.tile_w), ovx, ovy)
    h, v = (res[0]["dy"], res[0]["dx"]), (res[1]["dy"], res[1]["dx"])
    assert h == tuple(st.h_shift) and v == tuple(st.v_shift)
    lat = geo.Lattice(h, v)
    width, height = geo.canvas_size(st.tile_w, st.tile_h, xs, ys, st.pixel_size_um, lat)
    ctx.clear_fields()
    for c, ff in st.flatfields.items():
        ctx.set_flatfield(c, ff)
    job = []
    for t in tiles:                                  # paste order = sorted file names (fov 10 < fov 2)
        p = geo.place_tile(t.x_mm, t.y_mm, st.tile_w, st.tile_h, xs, ys, st.pixel_size_um, lat)
        job.append((t.pixels, p.x, p.y, st.monochrome_channels.index(t.channel), t.z_level, p.crop_t, p.crop_b,
                    p.crop_l, p.crop_r))
    out = np.zeros((1, 3, 1, height, width), np.uint16)
    ctx.fuse_region(job, (st.tile_h, st.tile_w), (3, 1, height, width), out=out, apply_flatfield=True)
    assert out.shape == exp.shape and np.array_equal(out, exp)
    ctx.clear_fields()


@pytest.mark.parametrize("mode", ["feather", "linear"])
def test_config3_2x2_flatfield_blend(ctx, mode):
    """2x2 tiles per well, flat-field + feathered (or linear) blend: extension modes, +-1 LSB against
    oracle/blend_ref.py (there is no reference behaviour for weighted blending)."""
    from image_stitcher_b200 import _ffi
    from image_stitcher_b200 import geometry as geo
    from oracle import blend_ref
    st, tiles, truth = synth.make_region(rows=2, cols=2, tile_h=512, tile_w=512, seed=63, jitter=2,
                                         channels=("Fluorescence 488 nm Ex", "Fluorescence 638 nm Ex"), apply_flatfield=True)
    xs = sorted(set(t.x_mm for t in tiles))
    ys = sorted(set(t.y_mm for t in tiles))
    width, height = geo.canvas_size(st.tile_w, st.tile_h, xs, ys, st.pixel_size_um, None)
    ov = geo.strip_overlaps(st.tile_w, st.tile_h, xs, ys, st.pixel_size_um, 2)
    ctx.clear_fields()
    flats = {c: ff.astype(np.float32) for c, ff in st.flatfields.items()}
    for c, ff in flats.items():
        ctx.set_flatfield(c, ff)
    job = []
    for t in tiles:
        p = geo.place_tile(t.x_mm, t.y_mm, st.tile_w, st.tile_h, xs, ys, st.pixel_size_um, None)
        job.append((t.pixels, p.x, p.y, st.monochrome_channels.index(t.channel), 0, 0, 0, 0, 0))
    out = np.zeros((1, 2, 1, height, width), np.uint16)
    ctx.fuse_region(job, (st.tile_h, st.tile_w), (2, 1, height, width), out=out, apply_flatfield=True,
                    blend=_ffi.BLEND_MODES[mode], blend_ov=ov)
    exp = blend_ref.fuse_blend(job, (2, 1, height, width), mode, ov=ov, flats=flats)
    diff = np.abs(out.astype(np.int32) - exp.astype(np.int32))
    assert diff.max() <= 1
    ctx.clear_fields()


def test_config4_zstack_planes_and_chunked_output(ctx):
    """5-plane z-stack into the zarr-chunk-ordered layout: every (c, z) plane lands in its own chunks."""
    from image_stitcher_b200 import _ffi
    from image_stitcher_b200 import geometry as geo
    st, tiles, truth = synth.make_region(rows=2, cols=3, tile_h=192, tile_w=256, seed=74, jitter=0, num_z=5)
    exp = sr.stitch_region(st, tiles)
    xs = sorted(set(t.x_mm for t in tiles))
    ys = sorted(set(t.y_mm for t in tiles))
    width, height = geo.canvas_size(st.tile_w, st.tile_h, xs, ys, st.pixel_size_um, None)
    job = []
    for t in tiles:
        p = geo.place_tile(t.x_mm, t.y_mm, st.tile_w, st.tile_h, xs, ys, st.pixel_size_um, None)
        job.append((t.pixels, p.x, p.y, 0, t.z_level, 0, 0, 0, 0))
    ch = 128
    ncy, ncx = -(-height // ch), -(-width // ch)
    out = np.full((5, ncy, ncx, ch, ch), 9, np.uint16)
    ctx.fuse_region(job, (st.tile_h, st.tile_w), (1, 5, height, width), out=out, layout=_ffi.SB_LAYOUT_CHUNKED,
                    chunk=(ch, ch))
    dense = out.transpose(0, 1, 3, 2, 4).reshape(5, ncy * ch, ncx * ch)
    assert np.array_equal(dense[:, :height, :width], exp[0, 0])
    assert not dense[:, height:, :].any() and not dense[:, :, width:].any()      # edge chunks zero padded


def test_config2_full_size_well_bit_exact_and_properties(ctx):
    """One well of BASELINE.json configs[2] at FULL size (3x3 tiles of 2048^2, 4 channels, flat-field on, coordinate
    placement -> (1, 4, 1, 5734, 5734)): bit-exact against the oracle, plus size-independent properties --
    idempotence (fusing twice gives the same canvas), and with unit flat-fields the canvas is a pure copy: every
    pixel equals the winning source pixel, so per-tile interiors hash to the same value as the inputs."""
    import hashlib
    from image_stitcher_b200 import _ffi
    from image_stitcher_b200 import geometry as geo
    chans = ("Fluorescence 405 nm Ex", "Fluorescence 488 nm Ex", "Fluorescence 561 nm Ex", "Fluorescence 638 nm Ex")
    st, tiles, truth = synth.make_region(rows=3, cols=3, tile_h=2048, tile_w=2048, seed=81, jitter=0, channels=chans,
                                         apply_flatfield=True)
    exp = sr.stitch_region(st, tiles)
    assert exp.shape == (1, 4, 1, 5734, 5734)
    xs = sorted(set(t.x_mm for t in tiles))
    ys = sorted(set(t.y_mm for t in tiles))
    job, origin = [], {}
    for t in tiles:
        p = geo.place_tile(t.x_mm, t.y_mm, 2048, 2048, xs, ys, st.pixel_size_um, None)
        c = st.monochrome_channels.index(t.channel)
        job.append((t.pixels, p.x, p.y, c, 0, 0, 0, 0, 0))
        origin[(t.fov, c)] = (p.x, p.y, t.pixels)
    ctx.clear_fields()
    for c, ff in st.flatfields.items():
        ctx.set_flatfield(c, ff)
    out = np.empty((1, 4, 1, 5734, 5734), np.uint16)
    ctx.fuse_region(job, (2048, 2048), (4, 1, 5734, 5734), out=out, apply_flatfield=True)
    assert np.array_equal(out, exp)
    again = np.empty_like(out)
    ctx.fuse_region(job, (2048, 2048), (4, 1, 5734, 5734), out=again, apply_flatfield=True)
    assert hashlib.sha256(again.tobytes()).digest() == hashlib.sha256(out.tobytes()).digest()
    # unit flat-field == no flat-field == copy of the winning pixels
    ctx.clear_fields()
    for c in range(4):
        ctx.set_flatfield(c, np.ones((2048, 2048), np.float32))
    unit = np.empty_like(out)
    ctx.fuse_region(job, (2048, 2048), (4, 1, 5734, 5734), out=unit, apply_flatfield=True)
    ctx.clear_fields()
    plain = np.empty_like(out)
    ctx.fuse_region(job, (2048, 2048), (4, 1, 5734, 5734), out=plain, apply_flatfield=False)
    assert np.array_equal(unit, plain)
    # the last tile in paste order (fov 8) is never overwritten: its whole footprint is a verbatim copy
    for c in range(4):
        x, y, px = origin[(8, c)]
        assert np.array_equal(plain[0, c, 0, y:y + 2048, x:x + 2048], px)
    # a tile's interior that no later tile touches (205-px overlaps) is a verbatim copy as well
    x, y, px = origin[(0, 2)]
    assert np.array_equal(plain[0, 2, 0, y:y + 1843, x:x + 1843], px[:1843, :1843])
