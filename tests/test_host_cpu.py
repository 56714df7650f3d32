"""CPU-side tests of the host mirror: parameters, Squid-layout parsing, placement geometry against the
reference-generated goldens, the OME-Zarr writer, the CLI flag surface, and that the C-ABI library
exports every symbol include/stitchb200.h declares (no compute calls: there is no GPU here)."""
import ctypes
import json
import os
import re

import numpy as np
import pytest

from conftest import RGB_GOLDENS, ROOT, SMALL_GOLDENS, load_golden
from image_stitcher_b200 import geometry as geo
from image_stitcher_b200 import ome_zarr_writer as ozw
from image_stitcher_b200.stitcher_parameters import StitchingParameters
from oracle import synth


# ------------------------------------------------------------------------------------------ C ABI
def _header_functions():
    src = open(os.path.join(ROOT, "include", "stitchb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sb_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_bound_exports():
    from image_stitcher_b200 import _ffi
    assert set(_header_functions()) == set(_ffi.EXPORTS)


def test_library_loads_and_exports_every_header_symbol():
    from image_stitcher_b200 import _ffi, build
    lib_path = build.build(force=False)
    lib = ctypes.CDLL(lib_path)
    missing = [s for s in _header_functions() if not hasattr(lib, s)]
    assert not missing, missing
    h = _ffi.load_library()
    assert h.sb_version() == 2
    assert h.sb_canvas_pitch(3891) == 3904 and h.sb_canvas_pitch(64) == 64
    assert h.sb_chunked_plane_elems(3891, 3891, 2048, 2048) == 4 * 2048 * 2048


def test_struct_layouts_match_the_header():
    """sizeof of the ctypes mirrors == what a C compiler gives for include/stitchb200.h."""
    import subprocess
    import tempfile
    from image_stitcher_b200 import _ffi
    code = ('#include "stitchb200.h"\n#include <stdio.h>\nint main(){printf("%zu %zu %zu %zu %zu\\n",sizeof(sb_tile),'
            'sizeof(sb_fuse_job),sizeof(sb_pair),sizeof(sb_pair_result),sizeof(sb_register_job));return 0;}')
    with tempfile.TemporaryDirectory() as tmp:
        c = os.path.join(tmp, "s.c")
        open(c, "w").write(code)
        exe = os.path.join(tmp, "s")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        sizes = [int(v) for v in subprocess.check_output([exe]).split()]
    mirrors = [_ffi.SbTile, _ffi.SbFuseJob, _ffi.SbPair, _ffi.SbPairResult, _ffi.SbRegisterJob]
    assert sizes == [ctypes.sizeof(m) for m in mirrors]


def test_no_device_means_loud_failure_not_a_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from image_stitcher_b200 import _ffi
    with pytest.raises(RuntimeError, match="sb_create"):
        _ffi.Context(0)


# ------------------------------------------------------------------------------------------ parameters / CLI
def test_parameters_roundtrip_and_validation(tmp_path):
    p = StitchingParameters(input_folder=str(tmp_path), use_registration=True, registration_channel="x",
                            apply_flatfield=True, scan_pattern="S-Pattern")
    p.validate()
    assert p.stitched_folder == p.stitched_folder            # stamp frozen (reference defect not copied)
    assert p.stitched_folder.startswith(str(tmp_path) + "_stitched_")
    j = tmp_path / "p.json"
    p.to_json(str(j))
    q = StitchingParameters.from_json(str(j))
    assert q.to_dict() == p.to_dict()
    assert StitchingParameters.from_dict({"input_folder": str(tmp_path), "unknown_key": 1}).output_format == ".ome.zarr"
    for bad in (dict(output_format=".png"), dict(scan_pattern="zigzag"), dict(blend_mode="max"),
                dict(use_registration=True, registration_z_level=-1), dict(upsample_factor=0)):
        with pytest.raises(ValueError):
            StitchingParameters(input_folder=str(tmp_path), **bad).validate()
    with pytest.raises(ValueError):
        StitchingParameters(input_folder=str(tmp_path / "missing")).validate()


def test_cli_flag_surface_matches_reference(tmp_path):
    from image_stitcher_b200 import stitcher_process_cli as cli
    a = cli.parse_args(["-i", str(tmp_path), "-r", "-ff", "-rc", "Fluorescence 488 nm Ex", "-rz", "1", "-s", "S-Pattern",
                        "-f", ".ome.zarr", "--dynamic-registration", "--merge-timepoints", "--merge-hcs-regions"])
    p = cli.create_params(a)
    assert (p.use_registration, p.apply_flatfield, p.registration_channel, p.registration_z_level) == \
        (True, True, "Fluorescence 488 nm Ex", 1)
    assert p.scan_pattern == "S-Pattern" and p.dynamic_registration and p.merge_timepoints and p.merge_hcs_regions
    assert p.blend_mode == "paste" and p.upsample_factor == 10          # extension defaults = reference behaviour
    pj = tmp_path / "params.json"
    json.dump({"input_folder": str(tmp_path), "use_registration": True}, open(pj, "w"))
    assert cli.create_params(cli.parse_args(["-i", "ignored", "--params-json", str(pj)])).use_registration


# ------------------------------------------------------------------------------------------ Squid layout parsing
def _stitcher(root, **params):
    from image_stitcher_b200.stitcher_process import StitcherProcess
    s = StitcherProcess(StitchingParameters(input_folder=root, **params), None, None, None, None)
    s.get_timepoints()
    s.extract_acquisition_parameters()
    s.get_pixel_size()
    s.parse_acquisition_metadata()
    return s


@pytest.mark.parametrize("name", SMALL_GOLDENS + RGB_GOLDENS)
def test_parse_squid_layout_and_geometry_match_reference_golden(name, tmp_path, capsys):
    g, st, tiles, kw = load_golden(name)
    root = str(tmp_path / "acq")
    synth.write_squid_layout(root, {"A1": tiles})
    s = _stitcher(root, use_registration=st.use_registration, apply_flatfield=st.apply_flatfield,
                  scan_pattern=st.scan_pattern, registration_channel=st.registration_channel)
    assert s.timepoints == ["0"] and s.regions == ["A1"]
    assert s.monochrome_channels == [str(c) for c in g["monochrome_channels"]]
    assert s.pixel_size_um == float(g["pixel_size_um"])
    assert (s.input_height, s.input_width) == (kw["tile_h"], kw["tile_w"]) and s.dtype == tiles[0].pixels.dtype
    assert s.num_z == kw.get("num_z", 1)
    # paste order == sorted file names == the order the reference iterated (golden tile_names)
    order = [os.path.basename(v["filepath"]) for v in s.get_region_data(0, "A1").values()]
    assert order == [str(n) for n in g["tile_names"]]
    # canvas size with the reference's registered shifts (incl. its height over-allocation quirk)
    s.h_shift, s.v_shift = tuple(int(v) for v in g["h_shift"]), tuple(int(v) for v in g["v_shift"])
    if st.scan_pattern == "S-Pattern":
        s.h_shift_rev = tuple(int(v) for v in g["h_shift_rev"])
        s.h_shift_rev_odd = int(g["h_shift_rev_odd"])
    w, h = s.calculate_output_dimensions(0, "A1")
    assert (1, s.num_c, s.num_z, h, w) == tuple(int(v) for v in g["canvas_shape"])


@pytest.mark.parametrize("name,world", [("reg_2x2_mono", 2), ("coord_2x2_plain", 3), ("reg_2x3_negdrift", 4), ("coord_2x3_rgb_u8", 2)])
def test_band_jobs_of_all_workers_paste_to_the_reference_canvas(name, world, tmp_path):
    """One region over several workers, host side only (no GPU): every worker's ``band_groups`` / ``_band_job`` --
    which files it decodes, how the tiles are re-based to the band and cropped -- pasted with plain NumPy slicing give,
    band by band, the rows of the reference's golden canvas (flat-field off here: a paste is then a copy).  The bands of
    all workers tile every plane exactly once."""
    g, st, tiles, kw = load_golden(name)
    if st.apply_flatfield:
        pytest.skip("paste == copy only without a flat-field")
    root = str(tmp_path / "acq")
    synth.write_squid_layout(root, {"A1": tiles})
    canvas = g["canvas"]
    seen = np.zeros(canvas.shape[1:4], dtype=int)                       # (C, Z, H) rows covered
    for rank in range(world):
        s = _stitcher(root, use_registration=st.use_registration, scan_pattern=st.scan_pattern,
                      registration_channel=st.registration_channel, rank=rank, world=world)
        s.chunks = (1, 1, 1, 64, 64)
        if st.use_registration:
            s.h_shift, s.v_shift = tuple(int(v) for v in g["h_shift"]), tuple(int(v) for v in g["v_shift"])
            if st.scan_pattern == "S-Pattern":
                s.h_shift_rev = tuple(int(v) for v in g["h_shift_rev"])
                s.h_shift_rev_odd = int(g["h_shift_rev_odd"])
        w, h = s.calculate_output_dimensions(0, "A1")
        assert (h, w) == canvas.shape[-2:]
        H, W = s.input_height, s.input_width
        for plane, y0, y1 in s.band_groups(h):
            job, c, z = s._band_job(0, "A1", plane, y0, y1)
            band = np.zeros((y1 - y0, w), canvas.dtype)
            for arr, x, y, cc, zz, ct, cb, cl, cr in job:
                assert (cc, zz) == (0, 0) and y + ct >= 0
                ys0, ys1 = y + ct, min(y + H - cb, y1 - y0)
                xs0, xs1 = max(x + cl, 0), min(x + W - cr, w)
                if ys1 > ys0 and xs1 > xs0:
                    band[ys0:ys1, xs0:xs1] = arr[ys0 - y:ys1 - y, xs0 - x:xs1 - x]
            assert np.array_equal(band, canvas[0, c, z, y0:y1]), (rank, plane, y0, y1)
            seen[c, z, y0:y1] += 1
    assert (seen == 1).all()


def test_parse_skips_hidden_and_focus_camera_and_handles_fov_ge_10(tmp_path):
    st, tiles, _ = synth.make_region(3, 4, 32, 32, seed=3, jitter=0, region="B2")
    root = str(tmp_path / "acq")
    synth.write_squid_layout(root, {"B2": tiles})
    import cv2
    cv2.imwrite(os.path.join(root, "0", ".hidden_0_0_x.tiff"), tiles[0].pixels)
    cv2.imwrite(os.path.join(root, "0", "B2_0_0_focus_camera.tiff"), tiles[0].pixels)
    s = _stitcher(root)
    names = [os.path.basename(v["filepath"]) for v in s.get_region_data(0, "B2").values()]
    assert len(names) == 12 and names == sorted(names)
    assert names.index("B2_10_0_Fluorescence_488_nm_Ex.tiff") < names.index("B2_2_0_Fluorescence_488_nm_Ex.tiff")
    with pytest.raises(ValueError):
        s.get_region_data(0, "nope")


def test_strip_overlap_rule_and_center_pairs():
    px = synth.pixel_size_um()
    xs = [10.0 + c * 1843 * px / 1000 for c in range(3)]
    ys = [20.0 + r * 1843 * px / 1000 for r in range(3)]
    assert geo.strip_overlaps(2048, 2048, xs, ys, px, 2) == (214, 214)      # round(205*1.05)//2*2
    assert geo.strip_overlaps(2048, 2048, xs, ys, px, 1) == (107, 107)      # the reference's binning-1 regime
    plan, rev_odd = geo.center_pairs(xs, ys, True)
    assert [k for k, _, _ in plan] == ["h", "v", "h_rev"] and rev_odd is False
    assert plan[0][1] == (xs[1], ys[1]) and plan[0][2] == (xs[2], ys[1]) and plan[1][2] == (xs[1], ys[2])
    assert len(geo.grid_pairs(3, 3)) == 12 and len(geo.grid_pairs(20, 20)) == 760


# ------------------------------------------------------------------------------------------ OME-Zarr writer
@pytest.mark.parametrize("compressor", [None, "zlib"])
def test_ome_zarr_roundtrip_and_metadata(tmp_path, compressor):
    rng = np.random.default_rng(0)
    data = rng.integers(0, 65535, (1, 2, 3, 150, 201), dtype=np.uint16)
    path = str(tmp_path / "A1_stitched.ome.zarr")
    ozw.write_ome_zarr(path, data, pixel_size_um=0.752, dz_um=1.5, channel_names=["a 405", "b 488"],
                       channel_colors=[0x0000FF, 0x00FF00], num_levels=3, chunks=(1, 1, 1, 64, 64), compressor=compressor)
    for l in range(3):
        assert np.array_equal(ozw.read_ome_zarr_level(path, l), data[..., ::2 ** l, ::2 ** l])
    attrs = json.load(open(os.path.join(path, ".zattrs")))
    ms = attrs["multiscales"][0]
    assert [a["name"] for a in ms["axes"]] == list("tczyx") and ms["version"] == "0.4"
    assert ms["datasets"][2]["coordinateTransformations"][0]["scale"] == [1, 1, 1.5, 0.752 * 4, 0.752 * 4]
    assert [c["color"] for c in attrs["omero"]["channels"]] == ["0000FF", "00FF00"]
    assert attrs["omero"]["channels"][0]["window"]["end"] == 65535
    za = json.load(open(os.path.join(path, "0", ".zarray")))
    assert za["chunks"] == [1, 1, 1, 64, 64] and za["dtype"] == "<u2" and za["dimension_separator"] == "/"
    # edge chunks are stored full size, zero padded (zarr v2)
    edge = os.path.join(path, "0", "0", "0", "0", "2", "3")
    if compressor is None:
        assert os.path.getsize(edge) == 64 * 64 * 2


def test_ome_zarr_from_chunk_ordered_buffer(tmp_path):
    rng = np.random.default_rng(1)
    C, Z, H, W, ch = 2, 1, 100, 130, 64
    dense = rng.integers(0, 65535, (C, Z, H, W), dtype=np.uint16)
    ncy, ncx = 2, 3
    buf = np.zeros((C * Z, ncy, ncx, ch, ch), np.uint16)
    for p in range(C * Z):
        for iy in range(ncy):
            for ix in range(ncx):
                blk = dense[p // Z, p % Z, iy * ch:(iy + 1) * ch, ix * ch:(ix + 1) * ch]
                buf[p, iy, ix, :blk.shape[0], :blk.shape[1]] = blk
    path = str(tmp_path / "c.ome.zarr")
    ozw.write_ome_zarr_chunked(path, buf, (C, Z, H, W), (ch, ch), pixel_size_um=1.0, channel_names=["a", "b"],
                               channel_colors=[1, 2])
    assert np.array_equal(ozw.read_ome_zarr_level(path, 0)[0], dense)


def test_ome_zarr_takes_precomputed_levels_and_pyramid_shapes(tmp_path):
    """``levels=`` (made by ``sb_pyramid`` on the GPU) are written as given; missing ones are sliced on the host."""
    from image_stitcher_b200 import _ffi
    rng = np.random.default_rng(2)
    data = rng.integers(0, 65535, (1, 2, 1, 101, 77), dtype=np.uint16)
    assert _ffi.Context.pyramid_shapes(data.shape, 4) == [(2, 51, 39), (2, 26, 20), (2, 13, 10)]
    assert _ffi.Context.pyramid_shapes(data.shape, 1) == []
    marked = data[..., ::2, ::2].copy()
    marked[0, 0, 0, 0, 0] ^= 1                        # prove the given level is the one stored
    path = str(tmp_path / "l.ome.zarr")
    ozw.write_ome_zarr(path, data, pixel_size_um=1.0, channel_names=["a", "b"], channel_colors=[1, 2], num_levels=3,
                       chunks=(1, 1, 1, 64, 64), levels=[marked])
    assert np.array_equal(ozw.read_ome_zarr_level(path, 1), marked)
    assert np.array_equal(ozw.read_ome_zarr_level(path, 2), marked[..., ::2, ::2])


@pytest.mark.parametrize("world", [1, 2, 3, 5])
def test_ome_zarr_bands_written_by_several_workers_equal_the_whole_store(tmp_path, world):
    """``write_ome_zarr_band``: the (plane, chunk-row) bands of ``shard.fusion_units_for_rank``, each written by its
    owner -- level 0 as whole chunks, coarser levels through memory maps of chunk files shared between workers (written
    here in REVERSE rank order) -- give the store ``write_ome_zarr`` writes from the whole canvas, level by level."""
    import json
    from image_stitcher_b200 import shard
    rng = np.random.default_rng(11)
    C, Z, H, W, ch, n_levels = 2, 2, 333, 205, 64, 4
    dense = rng.integers(1, 65535, (1, C, Z, H, W), dtype=np.uint16)
    whole, banded = str(tmp_path / "whole.ome.zarr"), str(tmp_path / "banded.ome.zarr")
    meta = dict(pixel_size_um=0.5, dz_um=1.5, channel_names=["a", "b"], channel_colors=[0xFF0000, 0x00FF00], name="R_t0")
    ozw.write_ome_zarr(whole, dense, num_levels=n_levels, chunks=(1, 1, 1, ch, ch), **meta)
    ncx = -(-W // ch)
    for rank in reversed(range(world)):
        for plane, y0, y1 in shard.fusion_units_for_rank(C * Z, H, ch, world, rank):
            c, z = divmod(plane, Z)
            band = dense[0, c, z, y0:y1]
            ncy = -(-band.shape[0] // ch)
            buf = np.zeros((1, ncy, ncx, ch, ch), np.uint16)
            for iy in range(ncy):
                for ix in range(ncx):
                    blk = band[iy * ch:(iy + 1) * ch, ix * ch:(ix + 1) * ch]
                    buf[0, iy, ix, :blk.shape[0], :blk.shape[1]] = blk
            levels, lv = [], band
            for _ in range(1, n_levels):
                lv = lv[::2, ::2]
                levels.append(lv[None, None, None])
            ozw.write_ome_zarr_band(banded, buf, levels, plane=(c, z), row0=y0, full_shape=(C, Z, H, W), chunk_hw=(ch, ch),
                                    n_levels=n_levels, **meta)
    for level in range(n_levels):
        a, b = ozw.read_ome_zarr_level(whole, level), ozw.read_ome_zarr_level(banded, level)
        assert a.shape == b.shape and np.array_equal(a, b), level
        assert json.load(open(os.path.join(whole, str(level), ".zarray"))) == json.load(open(os.path.join(banded, str(level), ".zarray")))
    assert json.load(open(os.path.join(whole, ".zattrs"))) == json.load(open(os.path.join(banded, ".zattrs")))
    with pytest.raises(ValueError):
        ozw.write_ome_zarr_band(banded, buf, [], plane=(0, 0), row0=ch + 1, full_shape=(C, Z, H, W), chunk_hw=(ch, ch),
                                n_levels=1, **meta)


def _band_writer_process(args):
    """One worker of the concurrent band-writer test (module level: it runs in a forked process)."""
    path, rank, world, seed, C, Z, H, W, ch, n_levels = args
    from image_stitcher_b200 import shard
    dense = np.random.default_rng(seed).integers(1, 65535, (1, C, Z, H, W), dtype=np.uint16)
    ncx = -(-W // ch)
    for plane, y0, y1 in shard.fusion_units_for_rank(C * Z, H, ch, world, rank):
        c, z = divmod(plane, Z)
        band = dense[0, c, z, y0:y1]
        ncy = -(-band.shape[0] // ch)
        buf = np.zeros((1, ncy, ncx, ch, ch), np.uint16)
        for iy in range(ncy):
            for ix in range(ncx):
                blk = band[iy * ch:(iy + 1) * ch, ix * ch:(ix + 1) * ch]
                buf[0, iy, ix, :blk.shape[0], :blk.shape[1]] = blk
        levels, lv = [], band
        for _ in range(1, n_levels):
            lv = lv[::2, ::2]
            levels.append(lv[None, None, None])
        ozw.write_ome_zarr_band(path, buf, levels, plane=(c, z), row0=y0, full_shape=(C, Z, H, W), chunk_hw=(ch, ch),
                                n_levels=n_levels, pixel_size_um=0.5, channel_names=["a", "b"], channel_colors=[1, 2], name="R_t0")
    return rank


def test_ome_zarr_bands_written_by_concurrent_processes(tmp_path):
    """Four worker PROCESSES write their bands into one store at the same time: level-0 chunk files have one owner, the
    coarser levels' chunk files are shared (each worker writes its rows through its own memory map), the metadata files
    are replaced atomically by everyone.  The result is the store of the whole canvas."""
    import subprocess
    import sys
    C, Z, H, W, ch, n_levels, seed, world = 2, 1, 1111, 517, 64, 5, 5, 4
    path = str(tmp_path / "shared.ome.zarr")
    code = ("import sys; sys.path[:0] = [%r, %r]; from test_host_cpu import _band_writer_process; "
            "_band_writer_process(eval(sys.argv[1]))" % (ROOT, os.path.dirname(os.path.abspath(__file__))))
    procs = [subprocess.Popen([sys.executable, "-c", code, repr((path, r, world, seed, C, Z, H, W, ch, n_levels))])
             for r in range(world)]                                   # all four at once (fresh interpreters: no fork of a threaded process)
    assert [p.wait(timeout=300) for p in procs] == [0] * world
    dense = np.random.default_rng(seed).integers(1, 65535, (1, C, Z, H, W), dtype=np.uint16)
    level = dense
    for l in range(n_levels):
        if l:
            level = level[..., ::2, ::2]
        assert np.array_equal(ozw.read_ome_zarr_level(path, l), level), l
    assert len(json.load(open(os.path.join(path, ".zattrs")))["multiscales"][0]["datasets"]) == n_levels
    assert not [f for f in os.listdir(path) if f.endswith(".tmp")]


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_visible_boxes_cover_exactly_what_paste_order_leaves_visible(seed):
    """``geometry.visible_boxes`` against a brute-force owner map: paste the tile INDEX of every kept rectangle in paste
    order; the pixels of tile i that survive lie inside its box, the box is tight in y and tight in x up to the 8-pixel
    rounding, and a tile with no surviving pixel gets ``None``.  Jittered 3 x 3 lattice with seam crops, two planes, one
    tile buried completely."""
    rng = np.random.default_rng(seed)
    th, tw, step = 48, 64, 40
    tiles = []
    for r in range(3):
        for c in range(3):
            for ch in range(2):
                x, y = c * step + int(rng.integers(0, 5)), r * step + int(rng.integers(0, 5))
                tiles.append((None, x, y, ch, 0, int(rng.integers(0, 3)), int(rng.integers(0, 3)), int(rng.integers(0, 3)),
                              int(rng.integers(0, 3))))
    tiles.insert(0, (None, 10, 10, 0, 0, 4, 4, 4, 4))            # buried under the first tiles of plane 0
    Hc, Wc = 2 * step + th + 2, 2 * step + tw + 1                 # clips the last row / column of tiles
    boxes = geo.visible_boxes(tiles, th, tw, Hc, Wc)
    owner = -np.ones((2, Hc, Wc), dtype=int)
    for i, (_, x, y, c, z, ct, cb, cl, cr) in enumerate(tiles):
        owner[c, max(y + ct, 0):min(y + th - cb, Hc), max(x + cl, 0):min(x + tw - cr, Wc)] = i
    assert boxes[0] is None and (owner != 0).all()
    for i, (t, box) in enumerate(zip(tiles, boxes)):
        ys, xs = np.nonzero(owner[t[3]] == i)
        if len(ys) == 0:
            assert box is None
            continue
        x0, y0, x1, y1 = box
        assert 0 <= x0 < x1 <= tw and 0 <= y0 < y1 <= th and x0 % 8 == 0 and (x1 % 8 == 0 or x1 == tw)
        assert y0 == ys.min() - t[2] and y1 == ys.max() + 1 - t[2]
        assert x0 <= xs.min() - t[1] < x0 + 8 and x1 - 8 < xs.max() + 1 - t[1] <= x1


def test_integration_md_stub_matches_the_binding():
    """The ctypes stub printed in INTEGRATION.md must stay in step with the header: its structures have the layout of
    the real binding's, and every prototype it sets names an exported symbol."""
    import ctypes as C
    import re
    from image_stitcher_b200 import _ffi
    text = open(os.path.join(os.path.dirname(__file__), "..", "INTEGRATION.md")).read()
    code = re.search(r"```python\n# stitcher_process_b200.py.*?```", text, re.S).group(0)
    structs = code[code.index("class SbPair("):code.index("lib.sb_create.argtypes")]
    ns = {"C": C}
    exec(structs, ns)
    for mine, real in [("SbPair", _ffi.SbPair), ("SbPairResult", _ffi.SbPairResult), ("SbRegisterJob", _ffi.SbRegisterJob),
                       ("SbTile", _ffi.SbTile), ("SbFuseJob", _ffi.SbFuseJob)]:
        assert C.sizeof(ns[mine]) == C.sizeof(real), mine
        assert [(n, C.sizeof(t)) for n, t in ns[mine]._fields_] == [(n, C.sizeof(t)) for n, t in real._fields_], mine
    for sym in re.findall(r"lib\.(sb_[a-z0-9_]+)", code):
        assert sym in _ffi.EXPORTS, sym


def test_get_flatfields_samples_like_the_reference(tmp_path, monkeypatch):
    """(:529-548) at most 32 random tiles per timepoint, stop once more than 48 are collected; one field per channel.
    The estimator itself is replaced by a recorder here (the CUDA estimator is covered by the GPU tests)."""
    import multiprocessing as mp
    import sys
    from image_stitcher_b200.stitcher_process import StitcherProcess
    monkeypatch.setitem(sys.modules, "basicpy", None)             # "not installed", whatever the environment holds
    chans = ("Fluorescence 405 nm Ex", "Fluorescence 488 nm Ex")
    st, tiles, _ = synth.make_region(5, 8, 16, 24, channels=chans, jitter=0)
    root = str(tmp_path / "acq")
    for t in range(3):
        synth.write_squid_layout(root, {"A1": tiles}, timepoint=t)
    s = StitcherProcess(StitchingParameters(input_folder=root, apply_flatfield=True), mp.Queue(), mp.Queue(), mp.Queue(), mp.Event())
    s.get_timepoints()
    s.extract_acquisition_parameters()
    s.get_pixel_size()
    s.parse_acquisition_metadata()
    calls = []

    class Recorder:
        def estimate_flatfield(self, sample, **kw):
            calls.append([np.asarray(t) for t in sample])
            return np.full(sample[0].shape, float(len(calls)), np.float32)

    s._ctx = Recorder()
    s.get_flatfields()
    s._ctx = None
    assert [len(c) for c in calls] == [64, 64]                    # 32 from timepoint 0, 32 more from timepoint 1, then > 48
    assert all(t.shape == (16, 24) and t.dtype == np.uint16 for c in calls for t in c)
    assert sorted(s.flatfields) == [0, 1] and float(s.flatfields[1][0, 0]) == 2.0
    by_channel = {ch: {t.pixels.tobytes() for t in tiles if t.channel == ch} for ch in chans}
    for ci, ch in enumerate(sorted(chans)):                        # every sampled tile belongs to the channel being fitted
        assert all(t.tobytes() in by_channel[ch] for t in calls[ci])
    from queue import Empty
    msgs = []
    while True:
        try:
            msgs.append(s.status_queue.get(timeout=0.5))
        except Empty:
            break
    assert any("not BaSiC" in str(m) for m in msgs)


def test_multi_device_workers_split_regions_without_exchange(tmp_path):
    """--devices 0,1: worker r of w stitches regions r, r + w, ...; both register the same first region."""
    from image_stitcher_b200 import stitcher_process_cli as cli
    from image_stitcher_b200.shard import wells_for_rank
    a = cli.parse_args(["-i", str(tmp_path), "--devices", "0,1,3"])
    assert [int(d) for d in a.devices.split(",")] == [0, 1, 3]
    regions = [f"{r}{c}" for r in "ABCD" for c in range(1, 7)]
    got = [[regions[i] for i in wells_for_rank(len(regions), 3, r)] for r in range(3)]
    assert sorted(sum(got, [])) == sorted(regions) and got[1][:2] == ["A2", "A5"]
    p = StitchingParameters(input_folder=str(tmp_path), rank=1, world=3, device=1)
    assert StitchingParameters.from_dict(p.to_dict()).world == 3


def test_solve_positions_recovers_per_tile_jitter():
    """geometry.solve_positions: consistent pairwise shifts of a jittered (non-lattice) grid -> exact tile origins."""
    rng = np.random.default_rng(0)
    R, C, W, H = 3, 4, 256, 192
    true = {(r, c): (c * 230 + int(rng.integers(-3, 4)), r * 172 + int(rng.integers(-3, 4))) for r in range(R) for c in range(C)}
    pairs, shifts = [], []
    for kind, (r0, c0), (r1, c1) in geo.grid_pairs(R, C):
        a, b = true[(r0, c0)], true[(r1, c1)]
        ddx, ddy = b[0] - a[0], b[1] - a[1]
        pairs.append((kind, r0 * C + c0, r1 * C + c1))
        shifts.append((ddy, ddx - W) if kind == "h" else (ddy - H, ddx))       # what calculate_*_shift returns
    pos = geo.solve_positions(R * C, W, H, pairs, shifts)
    mx, my = min(v[0] for v in true.values()), min(v[1] for v in true.values())
    assert all(pos[r * C + c] == (true[(r, c)][0] - mx, true[(r, c)][1] - my) for r in range(R) for c in range(C))
    assert geo.solve_positions(1, W, H, [], []) == [(0, 0)] and geo.solve_positions(0, W, H, [], []) == []
