"""GPU parity tests of the registration chain (K0-K4) through the C ABI.

Bar (BASELINE.json north_star): integer pixel shifts bit-exact to the reference's
skimage.phase_cross_correlation path; sub-pixel shifts within 1/upsample_factor."""
import hashlib

import numpy as np
import pytest

from conftest import GOLDEN_DIR, load_golden
from oracle import synth

pytestmark = pytest.mark.gpu

H_DIR, V_DIR = 0, 1


@pytest.fixture(scope="module")
def ctx():
    from image_stitcher_b200 import _ffi
    c = _ffi.Context(0)
    yield c
    c.close()


def oracle_pair(a, b, ov, direction, uf=10):
    from oracle import stitch_ref as sr
    fn = sr.calculate_horizontal_shift if direction == H_DIR else sr.calculate_vertical_shift
    return fn(a, b, ov, upsample_factor=uf, return_details=True)


@pytest.mark.parametrize("precision", [0, 1, 2])
def test_shift_calls_match_reference_golden(ctx, precision):
    g = np.load(f"{GOLDEN_DIR}/shift_calls.npz")
    for i in range(int(g["n"])):
        a, bh, bv, ov = g[f"a_{i}"], g[f"bh_{i}"], g[f"bv_{i}"], int(g[f"ov_{i}"])
        res = ctx.register_pairs([(a, bh, H_DIR), (a, bv, V_DIR)], a.shape, ov, ov, precision=precision)
        for r, key, other, d in ((res[0], f"h_{i}", bh, H_DIR), (res[1], f"v_{i}", bv, V_DIR)):
            (exp_int, exp_shift, det) = oracle_pair(a, other, ov, d)
            assert exp_int == tuple(g[key])                       # oracle == reference (sanity)
            assert (r["dy"], r["dx"]) == tuple(g[key]), (i, key, r, det)
            assert r["coarse"] == det["coarse"]
            assert np.abs(np.array(r["shift"]) - exp_shift).max() <= 0.1 + 1e-12
            assert r["ref_minmax"] == (int(a.min()), int(a.max()))
            assert r["mov_minmax"] == (int(other.min()), int(other.max()))


@pytest.mark.parametrize("name", ["reg_2x2_mono", "reg_3x3_spattern_flat", "reg_2x3_negdrift"])
def test_calculate_shifts_matches_reference_golden(ctx, name):
    from image_stitcher_b200 import geometry as geo
    g, st, tiles, kw = load_golden(name)
    xs = sorted(set(t.x_mm for t in tiles))
    ys = sorted(set(t.y_mm for t in tiles))
    ch = st.registration_channel or st.channel_names[0]
    ovx, ovy = geo.strip_overlaps(st.tile_w, st.tile_h, xs, ys, st.pixel_size_um, st.pixel_binning)
    pairs, rev_odd = geo.center_pairs(xs, ys, st.scan_pattern == "S-Pattern")
    lut = {(t.x_mm, t.y_mm): t.pixels for t in tiles if t.channel == ch and t.z_level == 0}
    job = [(lut[a], lut[b], H_DIR if kind.startswith("h") else V_DIR) for kind, a, b in pairs]
    res = ctx.register_pairs(job, (st.tile_h, st.tile_w), ovx, ovy)
    got = {kind: (r["dy"], r["dx"]) for (kind, _, _), r in zip(pairs, res)}
    assert got["h"] == tuple(g["h_shift"])
    assert got["v"] == tuple(g["v_shift"])
    if st.scan_pattern == "S-Pattern":
        assert got["h_rev"] == tuple(g["h_shift_rev"])
        assert int(rev_odd) == int(g["h_shift_rev_odd"])


def test_full_size_config0_bit_exact_and_truth(ctx):
    """BASELINE.json configs[0]: 2048^2 tiles -> strips 1024x214 / 214x1024 (214 = 2 * 107)."""
    from image_stitcher_b200 import geometry as geo
    from oracle import synth
    g, _, _, kw = load_golden("full_2x2_2048")
    st, tiles, truth = synth.make_region(**kw)
    if hashlib.sha256(np.stack([t.pixels for t in tiles]).tobytes()).hexdigest() != str(g["input_sha"]):
        pytest.skip("synthetic generator differs from the one that made the golden")
    xs = sorted(set(t.x_mm for t in tiles))
    ys = sorted(set(t.y_mm for t in tiles))
    ovx, ovy = geo.strip_overlaps(st.tile_w, st.tile_h, xs, ys, st.pixel_size_um, st.pixel_binning)
    assert (ovx, ovy) == (214, 214)
    lut = {(t.x_mm, t.y_mm): t.pixels for t in tiles}
    pairs, _ = geo.center_pairs(xs, ys, False)
    job = [(lut[a], lut[b], H_DIR if kind == "h" else V_DIR) for kind, a, b in pairs]
    for precision in (0, 1, 2):
        res = ctx.register_pairs(job, (2048, 2048), ovx, ovy, precision=precision)
        assert (res[0]["dy"], res[0]["dx"]) == tuple(g["h_shift"]) == truth["h_shift"]
        assert (res[1]["dy"], res[1]["dx"]) == tuple(g["v_shift"]) == truth["v_shift"]
    # indices and sub-pixel shift against the oracle
    for (kind, a, b), r in zip(pairs, res):
        exp_int, exp_shift, det = oracle_pair(lut[a], lut[b], 214, H_DIR if kind == "h" else V_DIR)
        assert r["coarse"] == det["coarse"] and r["fine"] == det["fine"]
        assert np.array_equal(np.array(r["shift"]), exp_shift)


@pytest.mark.parametrize("uf", [1, 4, 100])
def test_upsample_factors(ctx, uf):
    from oracle import synth
    st, tiles, truth = synth.make_region(rows=1, cols=2, tile_h=256, tile_w=512, seed=31, jitter=2)
    a, b = tiles[0].pixels, tiles[1].pixels
    ov = 56
    r = ctx.register_pairs([(a, b, H_DIR)], a.shape, ov, ov, upsample_factor=uf, precision=1)[0]
    exp_int, exp_shift, det = oracle_pair(a, b, ov, H_DIR, uf=uf)
    assert (r["dy"], r["dx"]) == exp_int
    assert r["coarse"] == det["coarse"]
    assert np.abs(np.array(r["shift"]) - exp_shift).max() <= 1.0 / uf + 1e-12


def test_low_signal_pair_falls_back_to_float64(ctx):
    """Pure-noise strips: the argmax is among near-equal values, AUTO must take the float64 path
    and then agree with the complex128 oracle."""
    rng = np.random.default_rng(3)
    a = rng.integers(0, 65536, (128, 192), dtype=np.uint16)
    b = rng.integers(0, 65536, (128, 192), dtype=np.uint16)
    r = ctx.register_pairs([(a, b, H_DIR), (a, b, V_DIR)], a.shape, 24, 20)
    for res, d, ov in ((r[0], H_DIR, 24), (r[1], V_DIR, 20)):
        assert res["precision"] == 1
        exp_int, exp_shift, det = oracle_pair(a, b, ov, d)
        assert (res["dy"], res["dx"]) == exp_int
        assert res["coarse"] == det["coarse"]


def test_constant_tile_and_zero_strip(ctx):
    """max == min is DEFINED as an all-zero normalised tile (the reference's NaN cast is undefined);
    a zero strip then gives P == 0 -> cc == 0 -> argmax index 0 in both the coarse and the fine stage."""
    a = np.full((64, 96), 1234, np.uint16)
    b = np.random.default_rng(1).integers(0, 4000, (64, 96), dtype=np.uint16)
    c = b.copy()
    c[:, -12:] = c.min()                      # non-constant tile whose strip is all zero after the stretch
    res = ctx.register_pairs([(a, b, H_DIR), (c, b, H_DIR), (b, a, V_DIR)], a.shape, 12, 10)
    for r, (x, y, d, ov) in zip(res, [(a, b, H_DIR, 12), (c, b, H_DIR, 12), (b, a, V_DIR, 10)]):
        exp_int, exp_shift, det = oracle_pair(x, y, ov, d)
        assert r["coarse"] == det["coarse"] == (0, 0) and r["fine"] == det["fine"] == (0, 0)
        assert (r["dy"], r["dx"]) == exp_int


def test_normalize_matches_oracle(ctx):
    from oracle import stitch_ref as sr
    rng = np.random.default_rng(5)
    tiles = np.stack([rng.integers(lo, hi, (96, 130), dtype=np.uint16)
                      for lo, hi in ((0, 65536), (100, 4000), (500, 501), (7, 8))])
    tiles[3][:] = 7
    out = ctx.normalize(tiles)
    for i in range(len(tiles)):
        assert np.array_equal(out[i], sr.normalize_image(tiles[i]))


def test_device_memory_batch(ctx):
    import torch
    from image_stitcher_b200 import _ffi
    from oracle import synth
    st, tiles, truth = synth.make_region(rows=2, cols=3, tile_h=256, tile_w=256, seed=41, jitter=2)
    grid = {(t.fov // 3, t.fov % 3): t.pixels for t in tiles}
    pool = torch.from_numpy(np.stack([grid[(r, c)] for r in range(2) for c in range(3)]).view(np.int16)).cuda()
    ptr = lambda r, c: pool[r * 3 + c].data_ptr()
    job, host_job = [], []
    from image_stitcher_b200 import geometry as geo
    for kind, (r0, c0), (r1, c1) in geo.grid_pairs(2, 3):
        d = H_DIR if kind == "h" else V_DIR
        job.append((ptr(r0, c0), ptr(r1, c1), d))
        host_job.append((grid[(r0, c0)], grid[(r1, c1)], d))
    torch.cuda.synchronize()
    res = ctx.register_pairs(job, (256, 256), 28, 28, mem=_ffi.SB_MEM_DEVICE)
    for r, (a, b, d) in zip(res, host_job):
        exp_int, _, det = oracle_pair(a, b, 28, d)
        assert (r["dy"], r["dx"]) == exp_int and r["coarse"] == det["coarse"]


def test_async_registration_matches_sync_and_overlaps_lanes(ctx):
    """sb_register_pairs_async: results land at sb_sync(lane); three lanes in flight at once give the same
    answers as the synchronous call (per-lane workspaces do not interfere)."""
    from image_stitcher_b200 import _ffi
    rng = np.random.default_rng(5)
    H, W, ov = 256, 320, 34
    jobs = []
    for lane in range(3):
        world = synth.make_world(H * 2 + 64, W * 2 + 64, rng)
        a = np.clip(world[20:20 + H, 20:20 + W], 0, 65535).astype(np.uint16)
        bh = np.clip(world[20 + lane:20 + lane + H, 20 + W - ov + 1:20 + 2 * W - ov + 1], 0, 65535).astype(np.uint16)
        bv = np.clip(world[20 + H - ov - lane:20 + 2 * H - ov - lane, 22:22 + W], 0, 65535).astype(np.uint16)
        jobs.append([(a, bh, _ffi.SB_DIR_HORIZONTAL), (a, bv, _ffi.SB_DIR_VERTICAL)])
    sync_res = [ctx.register_pairs(j, (H, W), ov, ov, lane=0) for j in jobs]
    pend = [ctx.register_pairs_async(j, (H, W), ov, ov, lane=lane) for lane, j in enumerate(jobs)]
    got = [p.get() for p in pend]
    for s, g in zip(sync_res, got):
        assert [(r["dy"], r["dx"], r["coarse"], r["fine"]) for r in s] == [(r["dy"], r["dx"], r["coarse"], r["fine"]) for r in g]
    # a second job on a lane with a parked one completes the parked one first
    p0 = ctx.register_pairs_async(jobs[0], (H, W), ov, ov, lane=1)
    again = ctx.register_pairs(jobs[1], (H, W), ov, ov, lane=1)
    assert [(r["dy"], r["dx"]) for r in again] == [(r["dy"], r["dx"]) for r in sync_res[1]]
    assert [(r["dy"], r["dx"]) for r in p0.get()] == [(r["dy"], r["dx"]) for r in sync_res[0]]
    assert ctx.register_pairs_async([], (H, W), ov, ov).get() == []


@pytest.mark.parametrize("partial_upload", [False, True])
def test_well_pipeline_host_buffers_match_oracle(ctx, partial_upload):
    """The end-to-end call the benchmark times (host tiles in -> shifts + host canvas out over 3 lanes); also with the
    upload restricted to the pixels that can reach the canvas (``partial_upload``)."""
    import torch
    from image_stitcher_b200 import _ffi
    from image_stitcher_b200.pipeline import WellPipeline
    from image_stitcher_b200.plate import PlateSpec, make_plate
    from oracle import stitch_ref as sr
    spec = PlateSpec(wells=5, rows=2, cols=2, tile_h=256, tile_w=256, channels=2, reg_channel=1, jitter=2, seed=3)
    plate = make_plate(spec, device="cuda:0")
    ctx.clear_fields()
    for c in range(spec.channels):
        ctx.set_flatfield(c, plate.flat[c], mem=_ffi.SB_MEM_DEVICE)
    pipe = WellPipeline(ctx, spec, apply_flatfield=True, partial_upload=partial_upload)
    # partial uploads: the staging buffers are reused by later wells, so a kernel that read a pixel outside the uploaded
    # boxes would see another well's data and fail the comparison below
    assert (pipe.uploads is not None and pipe.upload_bytes < pipe.well_bytes) if partial_upload else pipe.uploads is None
    Wc, Hc = spec.canvas_size()
    host = [plate.pool[w].cpu().numpy().view(np.uint16).copy() for w in range(spec.wells)]
    outs = [np.zeros((1, spec.channels, spec.num_z, Hc, Wc), np.uint16) for _ in range(spec.wells)]
    pend = [pipe.submit(host[w], outs[w]) for w in range(spec.wells)]
    pipe.drain()
    names = [f"ch{c}" for c in range(spec.channels)]
    xs, ys = spec.stage_positions()
    from image_stitcher_b200.plate import well_pairs
    ovx, ovy = spec.strip_overlaps()
    for w in range(spec.wells):
        hp, _ = well_pairs(spec, lambda r, c, ch, z: host[w][r, c, ch, z])
        exp = [(sr.calculate_horizontal_shift(a, b, ovx) if d == 0 else sr.calculate_vertical_shift(a, b, ovy))
               for a, b, d in hp]
        assert [(r["dy"], r["dx"]) for r in pend[w].get()] == exp
        st = sr.RegionState(tile_h=spec.tile_h, tile_w=spec.tile_w, pixel_size_um=spec.pixel_size_um,
                            monochrome_channels=names, channel_names=names, apply_flatfield=True,
                            flatfields={c: plate.flat[c].cpu().numpy() for c in range(spec.channels)})
        recs = []
        for fov in sorted(range(spec.rows * spec.cols), key=str):
            r, c = divmod(fov, spec.cols)
            for ch in range(spec.channels):
                recs.append(sr.TileRec(x_mm=xs[c], y_mm=ys[r], z_level=0, channel=names[ch], pixels=host[w][r, c, ch, 0], fov=fov))
        assert np.array_equal(outs[w], sr.stitch_region(st, recs))
    pipe.close()
    ctx.clear_fields()


def test_uint8_tiles_registration_matches_oracle(ctx):
    """8-bit tiles: normalize_image scales to iinfo(uint8).max = 255 before the phase correlation (:854)."""
    from oracle import stitch_ref as sr
    rng = np.random.default_rng(8)
    H, W, ov = 256, 320, 34
    world = synth.make_world(H * 2 + 64, W * 2 + 64, rng)
    a = (np.clip(world[20:20 + H, 20:20 + W], 0, 65535).astype(np.uint16) >> 8).astype(np.uint8)
    bh = (np.clip(world[22:22 + H, 20 + W - ov - 1:20 + 2 * W - ov - 1], 0, 65535).astype(np.uint16) >> 8).astype(np.uint8)
    bv = (np.clip(world[20 + H - ov + 1:20 + 2 * H - ov + 1, 17:17 + W], 0, 65535).astype(np.uint16) >> 8).astype(np.uint8)
    res = ctx.register_pairs([(a, bh, H_DIR), (a, bv, V_DIR)], (H, W), ov, ov)
    eh = sr.calculate_horizontal_shift(a, bh, ov, dtype=np.uint8, return_details=True)
    ev = sr.calculate_vertical_shift(a, bv, ov, dtype=np.uint8, return_details=True)
    for r, e in zip(res, (eh, ev)):
        assert (r["dy"], r["dx"]) == e[0]
        assert r["coarse"] == e[2]["coarse"] and r["fine"] == e[2]["fine"]
    pend = ctx.register_pairs_async([(a, bh, H_DIR)], (H, W), ov, ov, lane=2)
    assert (pend.get()[0]["dy"], pend.get()[0]["dx"]) == eh[0]


def test_full_range_and_exact_quotient_tiles(ctx):
    """normalize_image on tiles whose range makes quotients exact: max - min == 65535 (every pixel) and a range that
    divides 65535 * a for many a (b = 255 * 5) -- the integer stretch must agree with the float64 expression."""
    from oracle import stitch_ref as sr
    rng = np.random.default_rng(12)
    a = rng.integers(0, 65536, (64, 96), dtype=np.uint16)
    a[0, 0], a[0, 1] = 0, 65535
    b = rng.integers(1000, 1000 + 1276, (64, 96), dtype=np.uint16)
    b[0, 0], b[0, 1] = 1000, 1000 + 1275
    for t in (a, b):
        assert np.array_equal(ctx.normalize(t), sr.normalize_image(t))
    res = ctx.register_pairs([(a, b, H_DIR), (b, a, V_DIR)], a.shape, 12, 10, precision=1)
    for r, (x, y, d, ov) in zip(res, [(a, b, H_DIR, 12), (b, a, V_DIR, 10)]):
        exp_int, exp_shift, det = oracle_pair(x, y, ov, d)
        assert r["coarse"] == det["coarse"] and (r["dy"], r["dx"]) == exp_int
