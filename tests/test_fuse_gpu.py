"""GPU parity tests of the fusion kernel (K5) through the C ABI against the oracle and the
reference-generated golden canvases.  Bar: bit-exact for paste mode; within 1 LSB for the
(extension) blend modes."""
import numpy as np
import pytest

from conftest import SMALL_GOLDENS, load_golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from image_stitcher_b200 import _ffi
    c = _ffi.Context(0)
    yield c
    c.close()


def golden_job(g, st, tiles):
    """sb_tile tuples for a golden case using the PRODUCT geometry and the golden's shifts."""
    from image_stitcher_b200 import geometry as geo
    xs = sorted(set(t.x_mm for t in tiles))
    ys = sorted(set(t.y_mm for t in tiles))
    lat = None
    if st.use_registration:
        lat = geo.Lattice(tuple(int(v) for v in g["h_shift"]), tuple(int(v) for v in g["v_shift"]),
                          tuple(int(v) for v in g["h_shift_rev"]), int(g["h_shift_rev_odd"]),
                          st.scan_pattern == "S-Pattern")
    width, height = geo.canvas_size(st.tile_w, st.tile_h, xs, ys, st.pixel_size_um, lat)
    job = []
    for t in tiles:
        p = geo.place_tile(t.x_mm, t.y_mm, st.tile_w, st.tile_h, xs, ys, st.pixel_size_um, lat)
        job.append((np.ascontiguousarray(t.pixels), p.x, p.y, st.monochrome_channels.index(t.channel), t.z_level,
                    p.crop_t, p.crop_b, p.crop_l, p.crop_r))
    return job, (st.num_c, st.num_z, height, width)


@pytest.mark.parametrize("name", SMALL_GOLDENS)
def test_paste_matches_reference_golden(ctx, name):
    g, st, tiles, kw = load_golden(name)
    job, cshape = golden_job(g, st, tiles)
    assert (1,) + cshape == tuple(g["canvas_shape"])
    ctx.clear_fields()
    for c, ff in st.flatfields.items():
        ctx.set_flatfield(c, ff)
    out = np.full((1,) + cshape, 0xAB, g["canvas"].dtype)       # canvas dtype == tile dtype (uint16, or uint8)
    ctx.fuse_region(job, (st.tile_h, st.tile_w), cshape, out=out, apply_flatfield=st.apply_flatfield)
    assert np.array_equal(out, g["canvas"])


def random_job(rng, n_tiles, th, tw, C, Z, Hc, Wc, crops=True):
    job = []
    for i in range(n_tiles):
        px = rng.integers(0, 65536, (th, tw), dtype=np.uint16)
        x = int(rng.integers(0, max(1, Wc - tw // 2)))
        y = int(rng.integers(0, max(1, Hc - th // 2)))
        cr = [int(v) for v in rng.integers(0, 9, 4)] if crops else [0, 0, 0, 0]
        job.append((px, x, y, int(rng.integers(0, C)), int(rng.integers(0, Z)), *cr))
    return job


@pytest.mark.parametrize("seed,fields", [(0, "none"), (1, "flat"), (2, "flat+dark"), (3, "flat")])
def test_paste_random_geometry(ctx, seed, fields):
    from oracle import blend_ref
    rng = np.random.default_rng(seed)
    th, tw, C, Z, Hc, Wc = 96, 136, 2, 2, 333, 517
    job = random_job(rng, 23, th, tw, C, Z, Hc, Wc)
    flats = darks = None
    ctx.clear_fields()
    if "flat" in fields:
        flats = {0: rng.uniform(0.6, 1.4, (th, tw)).astype(np.float32)}          # channel 1 passes through
        if seed == 3:
            flats[0][5, 7] = 0.0                                                  # x/0 -> inf -> 65535 ; 0/0 -> 0
            job[0][0][5, 7] = 0
        ctx.set_flatfield(0, flats[0])
    if "dark" in fields:
        darks = {0: rng.uniform(0, 300, (th, tw)).astype(np.float32), 1: rng.uniform(0, 300, (th, tw)).astype(np.float32)}
        for c, d in darks.items():
            ctx.set_darkfield(c, d)
    out = np.full((1, C, Z, Hc, Wc), 7, np.uint16)
    ctx.fuse_region(job, (th, tw), (C, Z, Hc, Wc), out=out, apply_flatfield=fields != "none")
    exp = blend_ref.fuse_paste(job, (C, Z, Hc, Wc), flats, darks)
    assert np.array_equal(out, exp)


@pytest.mark.parametrize("mode", ["linear", "feather"])
@pytest.mark.parametrize("fields", ["none", "flat+dark"])
def test_blend_modes_within_one_lsb(ctx, mode, fields):
    from image_stitcher_b200 import _ffi
    from oracle import blend_ref
    rng = np.random.default_rng(5)
    th, tw, C, Z, Hc, Wc = 128, 160, 2, 1, 300, 420
    job = random_job(rng, 14, th, tw, C, Z, Hc, Wc, crops=False)
    flats = darks = None
    ctx.clear_fields()
    if fields != "none":
        flats = {c: rng.uniform(0.7, 1.2, (th, tw)).astype(np.float32) for c in range(C)}
        darks = {c: rng.uniform(0, 100, (th, tw)).astype(np.float32) for c in range(C)}
        for c in range(C):
            ctx.set_flatfield(c, flats[c])
            ctx.set_darkfield(c, darks[c])
    out = np.empty((1, C, Z, Hc, Wc), np.uint16)
    ov = (17, 13)
    ctx.fuse_region(job, (th, tw), (C, Z, Hc, Wc), out=out, apply_flatfield=fields != "none",
                    blend=_ffi.BLEND_MODES[mode], blend_ov=ov)
    exp = blend_ref.fuse_blend(job, (C, Z, Hc, Wc), mode, ov, flats, darks)
    diff = np.abs(out.astype(np.int32) - exp.astype(np.int32))
    assert diff.max() <= 1
    assert (diff != 0).mean() < 0.01           # disagreements only at exact .5 ties


@pytest.mark.parametrize("chunk,fields", [((128, 256), "none"), ((128, 256), "flat"), ((64, 384), "none"), ((128, 128), "flat+dark")])
def test_chunked_layout_equals_rowmajor(ctx, chunk, fields):
    """zarr-chunk-ordered output == row-major output, edge chunks zero padded.  Power-of-two chunk widths take the
    rectangle-streaming kernel, 384 and dark-fields the TMA kernel: the two kernels check each other."""
    from image_stitcher_b200 import _ffi
    rng = np.random.default_rng(8)
    th, tw, C, Z, Hc, Wc = 96, 120, 1, 2, 300, 411
    job = random_job(rng, 9, th, tw, C, Z, Hc, Wc)
    ctx.clear_fields()
    if "flat" in fields:
        ctx.set_flatfield(0, rng.uniform(0.6, 1.4, (th, tw)).astype(np.float32))
    if "dark" in fields:
        ctx.set_darkfield(0, rng.uniform(0, 300, (th, tw)).astype(np.float32))
    ref = np.empty((1, C, Z, Hc, Wc), np.uint16)
    ctx.fuse_region(job, (th, tw), (C, Z, Hc, Wc), out=ref, apply_flatfield=fields != "none")
    ch, cw = chunk
    ncy, ncx = -(-Hc // ch), -(-Wc // cw)
    out = np.full((C * Z, ncy, ncx, ch, cw), 3, np.uint16)
    ctx.fuse_region(job, (th, tw), (C, Z, Hc, Wc), out=out, layout=_ffi.SB_LAYOUT_CHUNKED, chunk=(ch, cw),
                    apply_flatfield=fields != "none")
    full = out.transpose(0, 1, 3, 2, 4).reshape(C * Z, ncy * ch, ncx * cw)
    assert np.array_equal(full[:, :Hc, :Wc], ref.reshape(C * Z, Hc, Wc))
    assert not full[:, Hc:, :].any() and not full[:, :, Wc:].any()      # zarr-v2 edge chunks are zero padded
    ctx.clear_fields()


def test_empty_region_is_zero(ctx):
    out = np.full((1, 1, 1, 70, 90), 5, np.uint16)
    ctx.fuse_region([], (64, 64), (1, 1, 70, 90), out=out)
    assert not out.any()


def test_device_memory_path_and_lanes(ctx):
    import torch
    from image_stitcher_b200 import _ffi
    from oracle import blend_ref
    rng = np.random.default_rng(9)
    th, tw, C, Z, Hc, Wc = 64, 128, 1, 1, 200, 500
    job = random_job(rng, 12, th, tw, C, Z, Hc, Wc)
    pool = torch.from_numpy(np.stack([t[0] for t in job]).view(np.int16)).cuda()
    dev_job = [(pool[i].data_ptr(),) + t[1:] for i, t in enumerate(job)]
    pitch = _ffi.canvas_pitch(Wc)
    outs = []
    ctx.clear_fields()
    for lane in range(ctx.num_lanes):
        out = torch.full((C * Z, Hc, pitch), -1, dtype=torch.int16, device="cuda")
        torch.cuda.synchronize()
        ctx.fuse_region(dev_job, (th, tw), (C, Z, Hc, Wc), out=out, tile_mem=_ffi.SB_MEM_DEVICE,
                        out_mem=_ffi.SB_MEM_DEVICE, lane=lane)
        outs.append(out)
    ctx.sync()
    exp = blend_ref.fuse_paste(job, (C, Z, Hc, Wc))
    for out in outs:
        got = out.cpu().numpy().view(np.uint16)
        assert np.array_equal(got[:, :, :Wc], exp.reshape(C * Z, Hc, Wc))
        assert not got[:, :, Wc:].any()


def test_errors_are_reported(ctx):
    with pytest.raises(RuntimeError, match="plane"):
        ctx.fuse_region([(np.zeros((8, 8), np.uint16), 0, 0, 3, 0, 0, 0, 0, 0)], (8, 8), (1, 1, 16, 16),
                        out=np.zeros((1, 1, 1, 16, 16), np.uint16))
    with pytest.raises(RuntimeError, match="negative"):
        ctx.fuse_region([(np.zeros((8, 8), np.uint16), -4, 0, 0, 0, 0, 0, 0, 0)], (8, 8), (1, 1, 16, 16),
                        out=np.zeros((1, 1, 1, 16, 16), np.uint16))


def test_band_sharded_fusion_equals_whole_canvas(ctx):
    """Multi-GPU partitioning of one big mosaic (shard.fusion_units_for_rank / tiles_for_band): fusing every
    (plane, chunk-row) band separately -- as ranks would -- reproduces the single-call canvas bit for bit."""
    from image_stitcher_b200 import shard
    rng = np.random.default_rng(11)
    th, tw, C, Z, Hc, Wc = 96, 136, 2, 1, 400, 517
    job = random_job(rng, 19, th, tw, C, Z, Hc, Wc)
    flat = rng.uniform(0.6, 1.4, (th, tw)).astype(np.float32)
    ctx.clear_fields()
    ctx.set_flatfield(1, flat)
    whole = np.zeros((1, C, Z, Hc, Wc), np.uint16)
    ctx.fuse_region(job, (th, tw), (C, Z, Hc, Wc), out=whole, apply_flatfield=True)
    parts = np.full_like(whole, 0x5555)
    world = 3
    for rank in range(world):
        for plane, y0, y1 in shard.fusion_units_for_rank(C * Z, Hc, 128, world, rank):
            c, z = divmod(plane, Z)
            sub = [t for t in job if t[3] == c and t[4] == z]
            band = [(t[0], t[1], t[2], 0, 0, *t[5:]) for t in shard.tiles_for_band(sub, th, y0, y1)]
            out = np.zeros((1, 1, 1, y1 - y0, Wc), np.uint16)
            ctx.clear_fields()
            if c == 1:
                ctx.set_flatfield(0, flat)
            ctx.fuse_region(band, (th, tw), (1, 1, y1 - y0, Wc), out=out, apply_flatfield=True)
            parts[0, c, z, y0:y1] = out[0, 0, 0]
    assert np.array_equal(parts, whole)


def test_uint8_pixels_paste_flatfield_and_elementwise(ctx):
    """SB_U8 (8-bit acquisitions): canvas dtype uint8, flat-field clip at 255 (stitcher_process.py:838-841), normalize_image
    scaled to 255 (:854) -- against plain NumPy statements of the reference lines."""
    rng = np.random.default_rng(21)
    th, tw, C, Z, Hc, Wc = 96, 136, 2, 1, 301, 407
    job = []
    for i in range(17):
        px = rng.integers(0, 256, (th, tw), dtype=np.uint8)
        job.append((px, int(rng.integers(0, Wc - 40)), int(rng.integers(0, Hc - 40)), int(rng.integers(0, C)), 0,
                    *[int(v) for v in rng.integers(0, 7, 4)]))
    flat = rng.uniform(0.5, 1.6, (th, tw)).astype(np.float32)
    ctx.clear_fields()
    ctx.set_flatfield(1, flat)
    for layout_chunked in (False, True):
        exp = np.zeros((1, C, Z, Hc, Wc), np.uint8)
        for px, x, y, c, z, ct, cb, cl, cr in job:
            v = px
            if c == 1:
                v = (px / flat).clip(min=0, max=255).astype(np.uint8)
            v = v[ct:th - cb, cl:tw - cr]
            y0, x0 = y + ct, x + cl
            y1, x1 = min(y0 + v.shape[0], Hc), min(x0 + v.shape[1], Wc)
            exp[0, c, z, y0:y1, x0:x1] = v[:y1 - y0, :x1 - x0]
        if not layout_chunked:
            out = np.full((1, C, Z, Hc, Wc), 77, np.uint8)
            ctx.fuse_region(job, (th, tw), (C, Z, Hc, Wc), out=out, apply_flatfield=True)
            assert out.dtype == np.uint8 and np.array_equal(out, exp)
        else:
            from image_stitcher_b200 import _ffi
            ch = 128
            ncy, ncx = -(-Hc // ch), -(-Wc // ch)
            out = np.full((C * Z, ncy, ncx, ch, ch), 77, np.uint8)
            ctx.fuse_region(job, (th, tw), (C, Z, Hc, Wc), out=out, apply_flatfield=True, layout=_ffi.SB_LAYOUT_CHUNKED,
                            chunk=(ch, ch))
            dense = out.transpose(0, 1, 3, 2, 4).reshape(C * Z, ncy * ch, ncx * ch)
            assert np.array_equal(dense[:, :Hc, :Wc], exp[0, :, 0])
    a = job[0][0]
    assert np.array_equal(ctx.flatfield_apply(1, a), (a / flat).clip(min=0, max=255).astype(np.uint8))
    from oracle import stitch_ref as sr
    assert np.array_equal(ctx.normalize(a), sr.normalize_image(a, np.uint8))
    with pytest.raises(RuntimeError, match="paste"):
        ctx.fuse_region(job, (th, tw), (C, Z, Hc, Wc), out=np.zeros((1, C, Z, Hc, Wc), np.uint8), blend=1, blend_ov=(10, 10))
    ctx.clear_fields()


@pytest.mark.parametrize("mode", ["linear", "feather"])
def test_blend_cells_path_flat_only_random_geometry(ctx, mode):
    """Blend modes with float32 flat-fields and no dark-field take the cell decomposition (single-cover cells through the
    rectangle-streaming kernel with rounding, overlap cells through blend_cells_kernel): +-1 LSB against the oracle,
    including crops, tiles hanging over the canvas edge and triple overlaps."""
    from image_stitcher_b200 import _ffi
    from oracle import blend_ref
    rng = np.random.default_rng(31)
    th, tw, C, Z, Hc, Wc = 96, 136, 2, 1, 300, 420
    job = []
    for i, (x, y) in enumerate([(0, 0), (110, 4), (220, 0), (300, 10), (6, 80), (118, 84), (230, 78), (0, 170), (125, 168),
                                (250, 172), (340, 215), (60, 40)]):
        px = rng.integers(0, 65536, (th, tw), dtype=np.uint16)
        job.append((px, x, y, i % C, 0, *[int(v) for v in rng.integers(0, 5, 4)]))
    flats = {0: rng.uniform(0.6, 1.4, (th, tw)).astype(np.float32)}
    ctx.clear_fields()
    ctx.set_flatfield(0, flats[0])
    out = np.full((1, C, Z, Hc, Wc), 9, np.uint16)
    ctx.fuse_region(job, (th, tw), (C, Z, Hc, Wc), out=out, apply_flatfield=True, blend=_ffi.BLEND_MODES[mode], blend_ov=(26, 16))
    exp = blend_ref.fuse_blend(job, (C, Z, Hc, Wc), mode, ov=(26, 16), flats=flats)
    diff = np.abs(out.astype(np.int32) - exp.astype(np.int32))
    assert diff.max() <= 1, (int(diff.max()), int((diff > 1).sum()))
    assert (diff > 0).mean() < 0.02                       # rounding ties only
    ctx.clear_fields()


def test_two_contexts_on_two_devices_in_one_process():
    """Every entry point binds the context's device itself: two contexts on different GPUs can be driven alternately
    from one thread without the caller switching devices (skipped on a single-GPU box)."""
    from image_stitcher_b200 import _ffi
    c0 = _ffi.Context(0)
    try:
        try:
            c1 = _ffi.Context(1)
        except RuntimeError:
            pytest.skip("needs two GPUs")
        try:
            rng = np.random.default_rng(3)
            H = W = 128
            tiles = [rng.integers(0, 65536, size=(H, W), dtype=np.uint16) for _ in range(4)]
            job = [(tiles[0], 0, 0, 0, 0, 0, 0, 0, 0), (tiles[1], 115, 2, 0, 0, 0, 0, 0, 0),
                   (tiles[2], 3, 117, 0, 0, 0, 0, 0, 0), (tiles[3], 118, 119, 0, 0, 0, 0, 0, 0)]
            shape = (1, 1, 1, 249, 247)
            flat = rng.uniform(0.7, 1.1, (H, W)).astype(np.float32)
            outs = []
            for c in (c0, c1, c0, c1):
                c.set_flatfield(0, flat)
                out = np.empty(shape, np.uint16)
                c.fuse_region(job, (H, W), shape[1:], out=out, apply_flatfield=True)
                lv = c.pyramid(shape, 2, dtype=_ffi.SB_U16)
                assert np.array_equal(lv[0], out[..., ::2, ::2])
                r = c.register_pairs([(tiles[0], tiles[0], _ffi.SB_DIR_HORIZONTAL)], (H, W), 16, 16)
                outs.append((out, (r[0]["dy"], r[0]["dx"])))
            for out, shift in outs[1:]:
                assert np.array_equal(out, outs[0][0]) and shift == outs[0][1]
        finally:
            c1.close()
    finally:
        c0.close()


# ------------------------------------------------------------------------------------------ sb_fuse_regions (batched)
def _plate_jobs(spec, plate, canvases, **kw):
    from image_stitcher_b200 import _ffi
    from image_stitcher_b200.plate import well_fuse_tiles
    Wc, Hc = spec.canvas_size()
    jobs = []
    for w in range(spec.wells):
        ptr = lambda r, c, ch, z, w=w: plate.pool[w, r, c, ch, z].data_ptr()
        jobs.append(dict(tiles=well_fuse_tiles(spec, ptr), tile_shape=(spec.tile_h, spec.tile_w),
                         canvas_shape=(spec.channels, spec.num_z, Hc, Wc), out=canvases[w], tile_mem=_ffi.SB_MEM_DEVICE,
                         out_mem=_ffi.SB_MEM_DEVICE, dtype=_ffi.SB_U16, **kw))
    return jobs


@pytest.mark.parametrize("layout", ["rowmajor", "chunked"])
def test_fuse_regions_batch_equals_region_by_region_and_oracle(ctx, layout):
    """One launch over all wells (channel-major) == one sb_fuse_region per well == the oracle, bit for bit."""
    import torch
    from image_stitcher_b200 import _ffi
    from image_stitcher_b200.plate import PlateSpec, make_plate
    from oracle import stitch_ref as sr
    spec = PlateSpec(wells=5, rows=3, cols=2, tile_h=192, tile_w=256, channels=3, num_z=2, jitter=2, seed=11)
    plate = make_plate(spec, device="cuda:0")
    ctx.clear_fields()
    for c in (0, 2):                                             # channel 1 has no field: passes through
        ctx.set_flatfield(c, plate.flat[c], mem=_ffi.SB_MEM_DEVICE)
    Wc, Hc = spec.canvas_size()
    planes = spec.channels * spec.num_z
    if layout == "rowmajor":
        pitch = _ffi.canvas_pitch(Wc)
        shape, kw = (spec.wells, planes, Hc, pitch), {}
    else:
        ch, cw = 128, 256
        ncy, ncx = -(-Hc // ch), -(-Wc // cw)
        shape, kw = (spec.wells, planes, ncy, ncx, ch, cw), dict(layout=_ffi.SB_LAYOUT_CHUNKED, chunk=(ch, cw))
    batch = torch.full(shape, 0x5A5A, dtype=torch.int16, device="cuda:0")
    single = torch.full(shape, 0x1111, dtype=torch.int16, device="cuda:0")
    torch.cuda.synchronize()
    launches0 = ctx.kernel_launches
    ctx.fuse_regions(_plate_jobs(spec, plate, batch, apply_flatfield=True, **kw))
    assert ctx.kernel_launches - launches0 == 1                  # the whole batch is one kernel launch
    for job in _plate_jobs(spec, plate, single, apply_flatfield=True, **kw):
        tiles = job.pop("tiles"); ts = job.pop("tile_shape"); cs = job.pop("canvas_shape")
        ctx.fuse_region(tiles, ts, cs, **job)
    assert torch.equal(batch, single)
    names = [f"ch{c}" for c in range(spec.channels)]
    xs, ys = spec.stage_positions()
    flats = {c: plate.flat[c].cpu().numpy() for c in (0, 2)}
    for w in (0, spec.wells - 1):
        host = plate.pool[w].cpu().numpy().view(np.uint16)
        st = sr.RegionState(tile_h=spec.tile_h, tile_w=spec.tile_w, pixel_size_um=spec.pixel_size_um, num_z=spec.num_z,
                            monochrome_channels=names, channel_names=names, apply_flatfield=True, flatfields=flats)
        recs = []
        for fov in sorted(range(spec.rows * spec.cols), key=str):
            r, c = divmod(fov, spec.cols)
            for z in range(spec.num_z):
                for chn in range(spec.channels):
                    recs.append(sr.TileRec(x_mm=xs[c], y_mm=ys[r], z_level=z, channel=names[chn], pixels=host[r, c, chn, z], fov=fov))
        exp = sr.stitch_region(st, recs)[0].reshape(planes, Hc, Wc)
        got = batch[w].cpu().numpy().view(np.uint16)
        if layout == "chunked":
            got = got.transpose(0, 1, 3, 2, 4).reshape(planes, ncy * ch, ncx * cw)
            assert not got[:, Hc:].any() and not got[:, :, Wc:].any()          # edge chunks are zero padded
        assert np.array_equal(got[:, :Hc, :Wc], exp)
    ctx.clear_fields()


@pytest.mark.parametrize("layout", ["rowmajor", "chunked"])
def test_fuse_regions_wide_tiles_equal_region_by_region(ctx, layout):
    """Tiles wide enough for the four-group interior chunks of ``paste_rect_jobs_kernel`` (flat-field vectors parked in
    shared memory, several regions per block), more regions than one block carries (8 + 1), every tile offset: the batch
    equals one ``sb_fuse_region`` per well (itself checked against the oracle above and in test_configs_gpu)."""
    import torch
    from image_stitcher_b200 import _ffi
    from image_stitcher_b200.plate import PlateSpec, make_plate
    spec = PlateSpec(wells=9, rows=2, cols=3, tile_h=96, tile_w=2560, channels=2, num_z=1, jitter=5, seed=23)
    plate = make_plate(spec, device="cuda:0")
    ctx.clear_fields()
    for c in range(spec.channels):
        ctx.set_flatfield(c, plate.flat[c], mem=_ffi.SB_MEM_DEVICE)
    Wc, Hc = spec.canvas_size()
    planes = spec.channels * spec.num_z
    if layout == "rowmajor":
        shape, kw = (spec.wells, planes, Hc, _ffi.canvas_pitch(Wc)), {}
    else:
        ch, cw = 64, 512
        shape, kw = (spec.wells, planes, -(-Hc // ch), -(-Wc // cw), ch, cw), dict(layout=_ffi.SB_LAYOUT_CHUNKED, chunk=(ch, cw))
    batch = torch.full(shape, 0x5A5A, dtype=torch.int16, device="cuda:0")
    single = torch.full(shape, 0x1111, dtype=torch.int16, device="cuda:0")
    torch.cuda.synchronize()
    ctx.fuse_regions(_plate_jobs(spec, plate, batch, apply_flatfield=True, **kw))
    for job in _plate_jobs(spec, plate, single, apply_flatfield=True, **kw):
        tiles = job.pop("tiles"); ts = job.pop("tile_shape"); cs = job.pop("canvas_shape")
        ctx.fuse_region(tiles, ts, cs, **job)
    assert torch.equal(batch, single)
    ctx.clear_fields()


@pytest.mark.parametrize("blend", ["linear", "feather"])
def test_fuse_regions_blend_batch_equals_region_by_region(ctx, blend):
    """Blend modes through ``sb_fuse_regions``: the cells are built once, the single-cover cells of every region go through
    one paste launch and the overlap cells through one ``blend_cells_kernel`` launch (region = grid z) -- bit-equal to one
    ``sb_fuse_region`` per well (itself within 1 LSB of oracle/blend_ref.py, test_blend_modes_*)."""
    import torch
    from image_stitcher_b200 import _ffi
    from image_stitcher_b200.plate import PlateSpec, make_plate
    spec = PlateSpec(wells=7, rows=2, cols=2, tile_h=256, tile_w=384, channels=3, num_z=1, jitter=2, seed=31)
    plate = make_plate(spec, device="cuda:0")
    ctx.clear_fields()
    for c in (0, 2):
        ctx.set_flatfield(c, plate.flat[c], mem=_ffi.SB_MEM_DEVICE)
    Wc, Hc = spec.canvas_size()
    ovx, ovy = spec.strip_overlaps()
    shape = (spec.wells, spec.channels * spec.num_z, Hc, _ffi.canvas_pitch(Wc))
    kw = dict(apply_flatfield=True, blend=_ffi.BLEND_MODES[blend], blend_ov=(ovx, ovy))
    batch = torch.full(shape, 0x5A5A, dtype=torch.int16, device="cuda:0")
    single = torch.full(shape, 0x1111, dtype=torch.int16, device="cuda:0")
    torch.cuda.synchronize()
    launches0 = ctx.kernel_launches
    ctx.fuse_regions(_plate_jobs(spec, plate, batch, **kw))
    assert ctx.kernel_launches - launches0 == 2                  # paste of the single-cover cells + blend of the overlaps
    for job in _plate_jobs(spec, plate, single, **kw):
        tiles = job.pop("tiles"); ts = job.pop("tile_shape"); cs = job.pop("canvas_shape")
        ctx.fuse_region(tiles, ts, cs, **job)
    assert torch.equal(batch, single)
    ctx.clear_fields()


def test_fuse_regions_falls_back_region_by_region(ctx):
    """Regions with different geometry (or host memory) are not batched: same results as separate calls."""
    rng = np.random.default_rng(21)
    th, tw, C, Z, Hc, Wc = 64, 96, 1, 1, 200, 260
    ctx.clear_fields()
    jobs, exp = [], []
    for k in range(3):
        job = random_job(rng, 7, th, tw, C, Z, Hc, Wc)
        out = np.zeros((1, C, Z, Hc, Wc), np.uint16)
        ref = np.zeros_like(out)
        ctx.fuse_region(job, (th, tw), (C, Z, Hc, Wc), out=ref)
        jobs.append(dict(tiles=job, tile_shape=(th, tw), canvas_shape=(C, Z, Hc, Wc), out=out))
        exp.append(ref)
    ctx.fuse_regions(jobs)
    for j, e in zip(jobs, exp):
        assert np.array_equal(j["out"], e)
    ctx.fuse_regions([])
