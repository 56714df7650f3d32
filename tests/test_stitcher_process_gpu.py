"""GPU parity of the reference-facing class: ``StitcherProcess`` of this package, driven the way the
reference's own harness drives its class (construct with queues, call the methods in ``run`` order,
inject flatfields by attribute) on the SAME on-disk Squid acquisitions that produced the goldens
(tests/golden/make_golden.py ran the unmodified reference on them).  Bit-exact shifts and canvases."""
import multiprocessing as mp
import os

import numpy as np
import pytest

from conftest import RGB_GOLDENS, SMALL_GOLDENS, load_golden
from image_stitcher_b200 import ome_zarr_writer as ozw
from image_stitcher_b200.stitcher_parameters import StitchingParameters
from oracle import synth

pytestmark = pytest.mark.gpu


def _make(root, st, **extra):
    from image_stitcher_b200.stitcher_process import StitcherProcess
    p = StitchingParameters(input_folder=root, use_registration=st.use_registration, apply_flatfield=st.apply_flatfield,
                            scan_pattern=st.scan_pattern, registration_channel=st.registration_channel, **extra)
    return StitcherProcess(p, mp.Queue(), mp.Queue(), mp.Queue(), mp.Event())


def _drain(q):
    from queue import Empty
    out = []
    while True:
        try:
            out.append(q.get(timeout=0.5))
        except Empty:
            return out


def _prepare(s):
    s.get_timepoints()
    s.extract_acquisition_parameters()
    s.get_pixel_size()
    s.parse_acquisition_metadata()


@pytest.mark.parametrize("name", SMALL_GOLDENS + RGB_GOLDENS)
def test_methods_in_run_order_match_reference_golden(name, tmp_path):
    g, st, tiles, kw = load_golden(name)
    root = str(tmp_path / "acq")
    synth.write_squid_layout(root, {"A1": tiles})
    s = _make(root, st)
    try:
        _prepare(s)
        if st.apply_flatfield:
            s.set_flatfields(st.flatfields)
        if st.use_registration:
            s.calculate_shifts(s.timepoints[0], s.regions[0])
            assert tuple(s.h_shift) == tuple(int(v) for v in g["h_shift"])
            assert tuple(s.v_shift) == tuple(int(v) for v in g["v_shift"])
            if st.scan_pattern == "S-Pattern":
                assert tuple(s.h_shift_rev) == tuple(int(v) for v in g["h_shift_rev"])
                assert int(s.h_shift_rev_odd) == int(g["h_shift_rev_odd"])
        out = s.stitch_region(0, "A1")
        assert out.shape == tuple(int(v) for v in g["canvas_shape"]) and out.dtype == g["canvas"].dtype
        assert np.array_equal(out, g["canvas"])
        assert "progress" in [m[0] for m in _drain(s.progress_queue)]
    finally:
        s.cleanup()


def test_single_tile_methods_match_oracle(tmp_path):
    """normalize_image / calculate_*_shift / apply_flatfield_correction / place_tile one call at a time."""
    from oracle import stitch_ref as sr
    g, st, tiles, kw = load_golden("reg_3x3_spattern_flat")
    root = str(tmp_path / "acq")
    synth.write_squid_layout(root, {"A1": tiles})
    s = _make(root, st)
    try:
        _prepare(s)
        s.set_flatfields(st.flatfields)
        a, b = tiles[0].pixels, tiles[1].pixels
        assert np.array_equal(s.normalize_image(a), sr.normalize_image(a))
        assert s.calculate_horizontal_shift(a, b, 14) == sr.calculate_horizontal_shift(a, b, 14)
        assert s.calculate_vertical_shift(a, b, 18) == sr.calculate_vertical_shift(a, b, 18)
        assert np.array_equal(s.apply_flatfield_correction(a, 1), sr.apply_flatfield_correction(st, a, 1))
        assert s.apply_flatfield_correction(a, 7) is a                       # no field for that channel: passthrough
        # per-tile placement API reproduces the batched canvas
        s.calculate_shifts(0, "A1")
        ref = s.stitch_region(0, "A1")
        canvas = s.init_output(0, "A1")
        xs, ys = list(s.x_positions), list(s.y_positions)
        from image_stitcher_b200 import geometry as geo
        for key, info in s.get_region_data(0, "A1").items():
            s.col_index, s.row_index = xs.index(info["x"]), ys.index(info["y"])
            p = geo.place_tile(info["x"], info["y"], s.input_width, s.input_height, xs, ys, s.pixel_size_um, s._lattice())
            from image_stitcher_b200.stitcher_process import read_image
            s.place_tile(canvas, read_image(info["filepath"]), p.x, p.y, key[3], key[4], 0)
        assert np.array_equal(canvas, ref)
    finally:
        s.cleanup()


def test_run_end_to_end_writes_ome_zarr(tmp_path):
    """BASELINE.json configs[0] in small: 2x2 grid, registration + stitch to OME-Zarr, through ``run()``."""
    g, st, tiles, kw = load_golden("reg_2x2_mono")
    root = str(tmp_path / "acq")
    synth.write_squid_layout(root, {"A1": tiles})
    s = _make(root, st)
    s.run()                                   # synchronously, like stitcher_cli.py:112
    kind, (path, dtype) = s.complete_queue.get(timeout=5)
    assert kind == "complete" and path.endswith("A1_stitched.ome.zarr") and os.path.isdir(path)
    assert np.array_equal(ozw.read_ome_zarr_level(path, 0), g["canvas"])
    import json
    levels = json.load(open(os.path.join(path, ".zattrs")))["multiscales"][0]["datasets"]
    assert len(levels) == s.num_pyramid_levels          # a7: max(1, ceil(log2(max(Wc, Hc) / 1024)))
    msgs = [m[1][0] for m in _drain(s.status_queue) if m[0] == "status"]
    assert any("Registration" in m for m in msgs) and any("Stitching" in m for m in msgs) and any("Saving" in m for m in msgs)


def test_dynamic_registration_all_pairs_median(tmp_path):
    """Extension behind the reference's unused flag: every adjacent pair in one batch, lower median per direction.
    Each pair's shift must equal the oracle's calculate_*_shift for that pair; the medians follow from them."""
    from image_stitcher_b200 import geometry as geo
    from oracle import stitch_ref as sr
    g, st, tiles, kw = load_golden("reg_2x3_negdrift")
    root = str(tmp_path / "acq")
    synth.write_squid_layout(root, {"A1": tiles})
    s = _make(root, st, dynamic_registration=True)
    try:
        _prepare(s)
        s.calculate_shifts(0, "A1")
        xs, ys = list(s.x_positions), list(s.y_positions)
        ovx, ovy = geo.strip_overlaps(s.input_width, s.input_height, xs, ys, s.pixel_size_um, s.pixel_binning)
        by_pos = {(t.x_mm, t.y_mm): t.pixels for t in tiles}
        exp = {"h": [], "v": []}
        for kind, (r0, c0), (r1, c1) in geo.grid_pairs(len(ys), len(xs)):
            a, b = by_pos[(xs[c0], ys[r0])], by_pos[(xs[c1], ys[r1])]
            exp[kind].append(sr.calculate_horizontal_shift(a, b, ovx) if kind == "h" else sr.calculate_vertical_shift(a, b, ovy))
        assert s.registration_results["h"] == exp["h"] and s.registration_results["v"] == exp["v"]
        assert len(exp["h"]) == 4 and len(exp["v"]) == 3

        def lower_median(lst, k):
            v = sorted(p[k] for p in lst)
            return v[(len(v) - 1) // 2]
        assert tuple(s.h_shift) == (lower_median(exp["h"], 0), lower_median(exp["h"], 1))
        assert tuple(s.v_shift) == (lower_median(exp["v"], 0), lower_median(exp["v"], 1))
    finally:
        s.cleanup()


def test_errors_flow_to_the_status_queue(tmp_path):
    g, st, tiles, kw = load_golden("coord_2x2_plain")
    root = str(tmp_path / "acq")
    synth.write_squid_layout(root, {"A1": tiles})
    s = _make(root, st)
    try:
        _prepare(s)
        with pytest.raises(ValueError):
            s.stitch_region(0, "Z9")
        errors = [m[1] for m in _drain(s.status_queue) if m[0] == "error"]
        assert errors and "Z9" in errors[0]
    finally:
        s.cleanup()


def test_stop_event_terminates_before_the_next_region(tmp_path):
    g, st, tiles, kw = load_golden("coord_2x2_plain")
    root = str(tmp_path / "acq")
    synth.write_squid_layout(root, {"A1": tiles})
    s = _make(root, st)
    _prepare(s)
    s.stop_event.set()
    with pytest.raises(SystemExit):
        s.stitch_region(0, "A1")


def test_stitcher_class_and_sync_cli(tmp_path, capsys):
    """The reference's second orchestrator (``stitcher.Stitcher``) and ``stitcher_cli``: same canvas, delivered through
    the Qt-style signals; the CLI runs in-process like stitcher_cli.py:112."""
    from image_stitcher_b200 import stitcher_cli
    from image_stitcher_b200.stitcher import Stitcher
    g, st, tiles, kw = load_golden("reg_2x2_mono")
    root = str(tmp_path / "acq")
    synth.write_squid_layout(root, {"A1": tiles})
    s = Stitcher(StitchingParameters(input_folder=root, use_registration=True))
    seen = {}
    s.finished_saving.connect(lambda path, dtype: seen.update(path=path, dtype=dtype))
    s.update_progress.connect(lambda cur, total: seen.update(progress=(cur, total)))
    s.run()
    assert s.chunks == (1, 1, 1, 512, 512) and seen["progress"] == (4, 4)
    assert np.array_equal(ozw.read_ome_zarr_level(seen["path"], 0), g["canvas"])
    with pytest.raises(ValueError):
        Stitcher(StitchingParameters(input_folder=str(tmp_path / "missing")))       # validate() in the constructor
    assert stitcher_cli.main(["-i", root, "-r"]) == 0
    assert "Stitching completed" in capsys.readouterr().out
    assert stitcher_cli.main(["-i", str(tmp_path / "missing")]) == 1


def _write_region(root, tiles):
    synth.write_squid_layout(root, {"A1": tiles})


def test_edge_cases_single_tile_single_row_and_ragged_grid(tmp_path):
    """Degenerate grids: 1x1 (nothing to register), 1x3 (no vertical pair -- the reference raises IndexError at :603, here
    the missing axis simply keeps its (0, 0) shift), and a ragged 2x2 grid whose centre-right tile is missing (the reference
    warns and keeps (0, 0), :630-633).  Canvases are compared with the oracle on the same tiles."""
    from oracle import stitch_ref as sr
    # ---- 1 x 1
    st, tiles, _ = synth.make_region(rows=1, cols=1, tile_h=96, tile_w=128, seed=91, jitter=0, use_registration=True)
    root = str(tmp_path / "one")
    _write_region(root, tiles)
    s = _make(root, st)
    try:
        _prepare(s)
        s.calculate_shifts(0, "A1")
        assert tuple(s.h_shift) == (0, 0) and tuple(s.v_shift) == (0, 0)
        out = s.stitch_region(0, "A1")
        assert out.shape == (1, 1, 1, 96, 128) and np.array_equal(out[0, 0, 0], tiles[0].pixels)
    finally:
        s.cleanup()
    # ---- 1 x 3 row with registration
    st, tiles, truth = synth.make_region(rows=1, cols=3, tile_h=192, tile_w=256, seed=92, jitter=2, use_registration=True)
    root = str(tmp_path / "row")
    _write_region(root, tiles)
    s = _make(root, st)
    try:
        _prepare(s)
        s.calculate_shifts(0, "A1")
        xs = sorted(set(t.x_mm for t in tiles))
        lut = {t.x_mm: t.pixels for t in tiles}
        ovx = geo_overlap(192, 256, xs, st.pixel_size_um)
        assert tuple(s.h_shift) == sr.calculate_horizontal_shift(lut[xs[1]], lut[xs[2]], ovx) and tuple(s.v_shift) == (0, 0)
        st.h_shift, st.v_shift = tuple(s.h_shift), (0, 0)
        assert np.array_equal(s.stitch_region(0, "A1"), sr.stitch_region(st, tiles))
    finally:
        s.cleanup()
    # ---- ragged 2 x 2: the tile right of the centre tile is missing
    st, tiles, _ = synth.make_region(rows=2, cols=2, tile_h=128, tile_w=160, seed=93, jitter=1, use_registration=True)
    xs, ys = sorted(set(t.x_mm for t in tiles)), sorted(set(t.y_mm for t in tiles))
    ragged = [t for t in tiles if not (t.x_mm == xs[1] and t.y_mm == ys[0])]
    root = str(tmp_path / "ragged")
    _write_region(root, ragged)
    s = _make(root, st)
    try:
        _prepare(s)
        s.calculate_shifts(0, "A1")
        assert tuple(s.h_shift) == (0, 0)                      # right neighbour of the centre tile is missing
        st2 = sr.calculate_shifts(st, ragged)
        assert tuple(s.v_shift) == tuple(st2.v_shift)
        assert np.array_equal(s.stitch_region(0, "A1"), sr.stitch_region(st2, ragged))
    finally:
        s.cleanup()


def geo_overlap(tile_h, tile_w, xs, px):
    from image_stitcher_b200 import geometry as geo
    return geo.strip_overlaps(tile_w, tile_h, xs, [0.0], px, 2)[0]


def test_global_placement_recovers_per_tile_jitter(tmp_path):
    """Extension ``placement='global'``: tiles cut from one noise-free world at lattice + independent per-tile jitter
    (which the reference's single-lattice model cannot express).  All adjacent pairs are registered on the GPU and the
    least-squares origins recover the true ones (within one pixel: a single pair may round the other way); the fused
    canvas is exactly the paste of the tiles at the solved origins."""
    from oracle import blend_ref
    from oracle.stitch_ref import TileRec
    rng = np.random.default_rng(77)
    R, C, H, W = 3, 3, 512, 512
    step_x, step_y, J = 460, 460, 2
    world = synth.make_world(R * step_y + H + 40, C * step_x + W + 40, rng).astype(np.float32)
    world = np.clip(world, 0, 65535).astype(np.uint16)
    px = synth.pixel_size_um()
    tiles, true = [], {}
    for r in range(R):
        for c in range(C):
            x = 10 + c * step_x + int(rng.integers(-J, J + 1))
            y = 10 + r * step_y + int(rng.integers(-J, J + 1))
            true[(r, c)] = (x, y)
            fov = r * C + c
            tiles.append(TileRec(x_mm=10.0 + c * step_x * px / 1000.0, y_mm=20.0 + r * step_y * px / 1000.0, z_level=0,
                                 channel="Fluorescence 488 nm Ex", pixels=world[y:y + H, x:x + W].copy(), fov=fov,
                                 name=f"A1_{fov}_0_Fluorescence_488_nm_Ex.tiff"))
    tiles.sort(key=lambda t: t.name)
    root = str(tmp_path / "acq")
    synth.write_squid_layout(root, {"A1": tiles})
    from image_stitcher_b200.stitcher_process import StitcherProcess
    s = StitcherProcess(StitchingParameters(input_folder=root, use_registration=True, placement="global"),
                        mp.Queue(), mp.Queue(), mp.Queue(), mp.Event())
    try:
        _prepare(s)
        pos = s.register_region_global(0, "A1")
        xs, ys = list(s.x_positions), list(s.y_positions)
        mx, my = min(v[0] for v in true.values()), min(v[1] for v in true.values())
        err = [max(abs(pos[(xs[c], ys[r])][0] - (true[(r, c)][0] - mx)), abs(pos[(xs[c], ys[r])][1] - (true[(r, c)][1] - my)))
               for r in range(R) for c in range(C)]
        assert max(err) <= 1 and sum(e == 0 for e in err) >= 6, err
        out = s.stitch_region(0, "A1")
        job = [(t.pixels, *pos[(t.x_mm, t.y_mm)], 0, 0, 0, 0, 0, 0) for t in tiles]      # paste order = sorted names
        assert np.array_equal(out, blend_ref.fuse_paste(job, (1, 1, out.shape[3], out.shape[4])))
    finally:
        s.cleanup()


def test_run_pipelines_several_regions_and_timepoints(tmp_path):
    """``run()`` over 3 regions x 2 timepoints: decode of the next region, fusion of the current one and the OME-Zarr
    write of the previous one overlap; every output equals the oracle's canvas of that (timepoint, region)."""
    import shutil
    from oracle import stitch_ref as sr
    root = str(tmp_path / "acq")
    regions = {}
    for k, name in enumerate(["A1", "A2", "B1"]):
        st, tiles, _ = synth.make_region(rows=2, cols=2, tile_h=128, tile_w=160, seed=100 + k, jitter=0, region=name)
        regions[name] = (st, tiles)
    synth.write_squid_layout(root, {n: t for n, (_, t) in regions.items()}, timepoint=0)
    shutil.copytree(os.path.join(root, "0"), os.path.join(root, "1"))
    s = _make(root, regions["A1"][0])
    s.run()
    kind, (path, dtype) = s.complete_queue.get(timeout=5)
    assert kind == "complete" and path.endswith(os.path.join("1_stitched", "B1_stitched.ome.zarr"))
    for t in (0, 1):
        for name, (st, tiles) in regions.items():
            p = os.path.join(s.output_folder, f"{t}_stitched", f"{name}_stitched.ome.zarr")
            assert np.array_equal(ozw.read_ome_zarr_level(p, 0), sr.stitch_region(st, tiles)), (t, name)
    saving = [m[1][0] for m in _drain(s.status_queue) if m[0] == "status" and "Saving" in m[1][0]]
    assert len(saving) == 6


def test_process_cli_worker_process_end_to_end(tmp_path):
    """``python -m image_stitcher_b200.stitcher_process_cli``: the parent never touches CUDA, the forked worker creates
    the context lazily, reports over the three queues and writes the OME-Zarr (the reference's CLI flow, :187-232)."""
    import subprocess
    import sys
    from conftest import ROOT
    g, st, tiles, kw = load_golden("reg_2x2_mono")
    root = str(tmp_path / "acq")
    synth.write_squid_layout(root, {"A1": tiles})
    r = subprocess.run([sys.executable, "-m", "image_stitcher_b200.stitcher_process_cli", "-i", root, "-r", "--devices", "0"],
                       cwd=ROOT, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "Stitching completed. Output saved to:" in r.stdout and "Status: Calculating Registration Shifts..." in r.stdout
    path = r.stdout.split("Output saved to:")[1].split()[0]
    assert np.array_equal(ozw.read_ome_zarr_level(path, 0), g["canvas"])
    bad = subprocess.run([sys.executable, "-m", "image_stitcher_b200.stitcher_process_cli", "-i", str(tmp_path / "missing")],
                         cwd=ROOT, capture_output=True, text=True, timeout=120)
    assert bad.returncode != 0


def test_visualize_registration_writes_the_reference_pngs(tmp_path):
    """a15: with visualize_registration the overlap strips of the registered pairs land in <out>/horizontal.png and
    vertical.png exactly as the reference's visualize_image builds them (:857-881); off by default."""
    import cv2
    from oracle import stitch_ref as sr
    g, st, tiles, kw = load_golden("reg_2x2_mono")
    root = str(tmp_path / "acq")
    synth.write_squid_layout(root, {"A1": tiles})
    s = _make(root, st, visualize_registration=True)
    try:
        _prepare(s)
        s.calculate_shifts(s.timepoints[0], s.regions[0])
        assert tuple(s.h_shift) == tuple(int(v) for v in g["h_shift"])
        from image_stitcher_b200 import geometry as geo
        xs, ys = list(s.x_positions), list(s.y_positions)
        ovx, ovy = geo.strip_overlaps(s.input_width, s.input_height, xs, ys, s.pixel_size_um, s.pixel_binning)
        lut = {(t.x_mm, t.y_mm): t.pixels for t in tiles}
        plan, _ = geo.center_pairs(xs, ys, False)
        for kind, a_xy, b_xy in plan:
            na, nb = sr.normalize_image(lut[a_xy]), sr.normalize_image(lut[b_xy])
            if kind == "h":
                m = int(na.shape[0] * 0.25)
                exp = np.hstack((na[m:-m, -ovx:], nb[m:-m, :ovx]))
                name = "horizontal.png"
            else:
                m = int(na.shape[1] * 0.25)
                exp = np.vstack((na[-ovy:, m:-m], nb[:ovy, m:-m]))
                name = "vertical.png"
            got = cv2.imread(os.path.join(s.output_folder, name), cv2.IMREAD_UNCHANGED)
            assert got is not None and np.array_equal(got, (exp / 65535 * 255).astype(np.uint8))
    finally:
        s.cleanup()
    s2 = _make(root, st)
    assert s2.visualize_registration is False


def test_tiles_of_another_shape_are_refused_before_the_c_abi(tmp_path):
    """a tile whose shape differs from (input_height, input_width) must not be handed to the library."""
    from image_stitcher_b200 import _ffi
    g, st, tiles, kw = load_golden("reg_2x2_mono")
    root = str(tmp_path / "acq")
    synth.write_squid_layout(root, {"A1": tiles})
    s = _make(root, st)
    try:
        _prepare(s)
        a = tiles[0].pixels
        with pytest.raises(ValueError):
            s.calculate_horizontal_shift(a, a[:-8], 14)
        with pytest.raises(ValueError):
            s.ctx.register_pairs([(a, np.ascontiguousarray(a[:, :-8]), 0)], a.shape, 14, 14)
        with pytest.raises(ValueError):
            s.ctx.fuse_region([(a[::2], 0, 0, 0, 0, 0, 0, 0, 0)], a.shape, (1, 1, 64, 64), out=np.zeros((1, 1, 1, 64, 64), np.uint16))
    finally:
        s.cleanup()


def test_run_fast_path_equals_the_plain_path_level_by_level(tmp_path, monkeypatch):
    """``run()`` on uint16 / paste / OME-Zarr jobs takes the pinned, chunk-ordered fast path (RegionPipeline): level 0
    arrives in zarr-chunk order from the GPU, levels 1.. from ``sb_pyramid`` on the resident canvas.  Every level of every
    region equals what the plain path (row-major canvas, host re-tiling; SB_NO_FAST_IO=1) writes, and level 0 equals the
    oracle; the multiscales / omero name follows the reference (f"{region}_t{timepoint}", :1086)."""
    import json
    import shutil
    from oracle import stitch_ref as sr
    root = str(tmp_path / "acq")
    regions = {}
    for k, name in enumerate(["A1", "A2", "A3", "A4"]):                    # more regions than lanes: buffers get reused
        st, tiles, _ = synth.make_region(rows=2, cols=3, tile_h=640, tile_w=768, seed=200 + k, jitter=0, region=name,
                                         channels=("Fluorescence 405 nm Ex", "Fluorescence 488 nm Ex"), apply_flatfield=True)
        regions[name] = (st, tiles)
    synth.write_squid_layout(root, {n: t for n, (_, t) in regions.items()}, timepoint=0)
    st0 = regions["A1"][0]

    def run_once(out_name, fast):
        if fast:
            monkeypatch.delenv("SB_NO_FAST_IO", raising=False)
        else:
            monkeypatch.setenv("SB_NO_FAST_IO", "1")
        s = _make(root, st0)
        s.set_flatfields(st0.flatfields)
        s.chunks = (1, 1, 1, 512, 512)                                     # several chunks per plane, ragged edges
        s.run()
        used = s.fast_io_used
        dst = str(tmp_path / out_name)
        shutil.move(s.output_folder, dst)
        return dst, used, s.num_pyramid_levels

    fast_dir, used_fast, n_levels = run_once("fast", True)
    plain_dir, used_plain, _ = run_once("plain", False)
    assert used_fast and not used_plain and n_levels >= 2
    for name, (st, tiles) in regions.items():
        pf = os.path.join(fast_dir, "0_stitched", f"{name}_stitched.ome.zarr")
        pp = os.path.join(plain_dir, "0_stitched", f"{name}_stitched.ome.zarr")
        exp = sr.stitch_region(st, tiles)
        for level in range(n_levels):
            a, b = ozw.read_ome_zarr_level(pf, level), ozw.read_ome_zarr_level(pp, level)
            assert a.shape == b.shape and np.array_equal(a, b), (name, level)
            if level == 0:
                assert np.array_equal(a, exp)
        za, zb = json.load(open(os.path.join(pf, ".zattrs"))), json.load(open(os.path.join(pp, ".zattrs")))
        assert za == zb and za["multiscales"][0]["name"] == f"{name}_t0" == za["omero"]["name"]


@pytest.mark.parametrize("world", [2, 3])
def test_one_region_split_over_several_workers_equals_the_single_worker_zarr(tmp_path, world):
    """SURVEY 8e inside the orchestrator: with fewer regions than workers every worker fuses ITS (plane, chunk-row)
    bands of the region (only the tiles that reach them are decoded) and drops the chunks into one shared OME-Zarr --
    level 0 chunk files have one owner, rows of the coarser levels go through memory maps of shared chunk files.  The
    result equals the single-worker store level by level, byte for byte, and level 0 equals the oracle.  (Workers run
    one after the other here, all on device 0: there is nothing to exchange, so order and concurrency cannot matter.)"""
    import json
    import shutil
    from oracle import stitch_ref as sr
    root = str(tmp_path / "acq")
    st, tiles, _ = synth.make_region(rows=3, cols=3, tile_h=640, tile_w=768, seed=77, jitter=2, region="B2",
                                     channels=("Fluorescence 405 nm Ex", "Fluorescence 488 nm Ex"), apply_flatfield=True,
                                     use_registration=True, registration_channel="Fluorescence 488 nm Ex")
    synth.write_squid_layout(root, {"B2": tiles}, timepoint=0)

    def worker(rank, n, stamp):
        s = _make(root, st, rank=rank, world=n)
        s.params._stamp = stamp
        s.output_folder = s.params.stitched_folder
        s.per_timepoint_region_output_template = os.path.join(s.output_folder, "{timepoint}_stitched",
                                                              "{region}_stitched" + s.output_format)
        s.set_flatfields(st.flatfields)
        s.chunks = (1, 1, 1, 512, 512)
        s.run()
        return s

    single = worker(0, 1, "single")
    assert single.fast_io_used and not single.band_mode
    ranks = [worker(r, world, f"split{world}") for r in range(world)]
    assert all(s.band_mode for s in ranks)
    n_levels = single.num_pyramid_levels
    assert n_levels >= 2
    ps = os.path.join(single.output_folder, "0_stitched", "B2_stitched.ome.zarr")
    pb = os.path.join(ranks[0].output_folder, "0_stitched", "B2_stitched.ome.zarr")
    exp = sr.stitch_region(st, tiles) if not st.use_registration else None
    for level in range(n_levels):
        a, b = ozw.read_ome_zarr_level(ps, level), ozw.read_ome_zarr_level(pb, level)
        assert a.shape == b.shape and np.array_equal(a, b), level
    if exp is not None:
        assert np.array_equal(ozw.read_ome_zarr_level(pb, 0), exp)
    assert json.load(open(os.path.join(ps, ".zattrs"))) == json.load(open(os.path.join(pb, ".zattrs")))
    for level in range(n_levels):
        assert json.load(open(os.path.join(ps, str(level), ".zarray"))) == json.load(open(os.path.join(pb, str(level), ".zarray")))
    # the bands of the workers are disjoint and cover the canvas
    H = ozw.read_ome_zarr_level(ps, 0).shape[-2]
    cover = np.zeros((single.num_c * single.num_z, -(-H // 512)), dtype=int)
    for s in ranks:
        for p, y0, y1 in s.band_groups(H):
            cover[p, y0 // 512:-(-y1 // 512)] += 1
    assert (cover == 1).all()
    shutil.rmtree(single.output_folder), shutil.rmtree(ranks[0].output_folder)


def test_process_cli_splits_one_region_over_two_concurrent_workers(tmp_path):
    """``--devices 0,0``: two worker PROCESSES at once (same GPU here; any two GPUs in production) share the one region
    of the acquisition by (plane, chunk-row) bands and write one OME-Zarr between them: level 0 equals the reference
    golden, the coarser levels (if the canvas is large enough to have any) equal ``[::2, ::2]`` of it."""
    import subprocess
    import sys
    from conftest import ROOT
    g, st, tiles, kw = load_golden("reg_2x2_mono")
    root = str(tmp_path / "acq")
    synth.write_squid_layout(root, {"A1": tiles})
    r = subprocess.run([sys.executable, "-m", "image_stitcher_b200.stitcher_process_cli", "-i", root, "-r", "--devices", "0,0"],
                       cwd=ROOT, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "of worker 0/2" in r.stdout and "of worker 1/2" in r.stdout
    path = r.stdout.split("Output saved to:")[1].split()[0]
    level = ozw.read_ome_zarr_level(path, 0)
    assert np.array_equal(level, g["canvas"])
    l = 1
    while os.path.isdir(os.path.join(path, str(l))):
        level = level[..., ::2, ::2]
        assert np.array_equal(ozw.read_ome_zarr_level(path, l), level), l
        l += 1
    import json
    assert l == len(json.load(open(os.path.join(path, ".zattrs")))["multiscales"][0]["datasets"])
