#!/usr/bin/env python3
"""Generate tests/golden/*.npz by executing the UNMODIFIED reference under import shims.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

For each case a synthetic Squid acquisition is written to a temp dir with
``oracle.synth``, the reference's own ``StitcherProcess`` parses it, runs
``calculate_shifts`` + ``stitch_region`` (``oracle.ref_shim``) and the inputs and
outputs are stored.  Small cases store the tiles and the canvas verbatim; the
full-size case (BASELINE.json configs[0]) stores the generator seed, a SHA-256 of
the regenerated inputs and a SHA-256 of the reference canvas.
"""
from __future__ import annotations

import hashlib
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_shim, synth  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))

CASES = {
    # name: kwargs for synth.make_region (+ params for the reference)
    "reg_2x2_mono": dict(rows=2, cols=2, tile_h=192, tile_w=256, seed=11, jitter=2,
                         use_registration=True),
    "reg_3x3_spattern_flat": dict(rows=3, cols=3, tile_h=160, tile_w=128, seed=12, jitter=2, num_z=2,
                                  channels=("Fluorescence 488 nm Ex", "Fluorescence 561 nm Ex"),
                                  use_registration=True, apply_flatfield=True, scan_pattern="S-Pattern",
                                  registration_channel="Fluorescence 561 nm Ex"),
    "coord_3x4_flat64": dict(rows=3, cols=4, tile_h=96, tile_w=128, seed=13, jitter=0,
                             channels=("BF LED matrix full", "Fluorescence 405 nm Ex"),
                             apply_flatfield=True),
    "coord_2x2_plain": dict(rows=2, cols=2, tile_h=128, tile_w=128, seed=14, jitter=0),
    "reg_2x3_negdrift": dict(rows=2, cols=3, tile_h=160, tile_w=192, seed=21, jitter=3, use_registration=True),
    # 8-bit acquisition (the reference takes every range from the dtype of the first image, :340/:838/:854)
    "reg_2x2_u8_flat": dict(rows=2, cols=2, tile_h=192, tile_w=256, seed=15, jitter=2, use_registration=True,
                            apply_flatfield=True, bits=8),
    # 8-bit RGB camera tiles: one file per fov, expanded to <channel>_R/_G/_B planes (:355-362, :757-763)
    "coord_2x3_rgb_u8": dict(rows=2, cols=3, tile_h=96, tile_w=128, seed=16, jitter=0, bits=8, rgb=True,
                             channels=("c0", "c1", "c2")),
}
FULL = {
    "full_2x2_2048": dict(rows=2, cols=2, tile_h=2048, tile_w=2048, seed=7, jitter=3, use_registration=True),
}


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def run_case(name, kw, store_arrays=True):
    gen_kw = {k: v for k, v in kw.items() if k not in ("bits", "rgb")}
    st, tiles, truth = synth.make_region(**gen_kw)
    if kw.get("bits") == 8:
        for t in tiles:
            t.pixels = (t.pixels >> 8).astype(np.uint8)
    if kw.get("rgb"):
        # merge the three synthetic channels of every (fov, z) into one H x W x 3 colour tile of channel "BF LED matrix full"
        by_key = {}
        for t in tiles:
            by_key.setdefault((t.fov, t.z_level), {})[t.channel] = t
        merged = []
        for (fov, z), chans in sorted(by_key.items()):
            first = chans["c0"]
            first.pixels = np.stack([chans[c].pixels for c in ("c0", "c1", "c2")], axis=-1)
            first.channel = "BF LED matrix full"
            first.name = f"A1_{fov}_{z}_BF_LED_matrix_full.tiff"
            merged.append(first)
        tiles = sorted(merged, key=lambda t: t.name)
    flat64 = name.endswith("flat64")
    with tempfile.TemporaryDirectory() as tmp:
        root = os.path.join(tmp, "acq")
        synth.write_squid_layout(root, {"A1": tiles})
        s = ref_shim.make_reference_stitcher(
            root, os.path.join(tmp, "out"),
            use_registration=st.use_registration, apply_flatfield=st.apply_flatfield,
            scan_pattern=st.scan_pattern, registration_channel=st.registration_channel)
        flatfields = None
        if st.apply_flatfield:
            flatfields = {c: (ff.astype(np.float64) if flat64 else ff) for c, ff in st.flatfields.items()}
        out = ref_shim.run_reference(s, flatfields=flatfields)
        canvas = out[(0, "A1")]
        rec = {
            "kwargs": np.array(repr(kw)),
            "h_shift": np.array(s.h_shift, dtype=np.int64),
            "v_shift": np.array(s.v_shift, dtype=np.int64),
            "h_shift_rev": np.array(getattr(s, "h_shift_rev", (0, 0)), dtype=np.int64),
            "h_shift_rev_odd": np.array(int(getattr(s, "h_shift_rev_odd", 0))),
            "canvas_shape": np.array(canvas.shape, dtype=np.int64),
            "canvas_sha": np.array(sha(canvas)),
            "truth_h": np.array(truth["h_shift"], dtype=np.int64),
            "truth_v": np.array(truth["v_shift"], dtype=np.int64),
            "pixel_size_um": np.array(s.pixel_size_um),
            "monochrome_channels": np.array(s.monochrome_channels),
            "input_sha": np.array(sha(np.stack([t.pixels for t in tiles]))),
            "tile_names": np.array([t.name for t in tiles]),
        }
        if store_arrays:
            rec["tiles"] = np.stack([t.pixels for t in tiles])
            rec["tile_x_mm"] = np.array([t.x_mm for t in tiles])
            rec["tile_y_mm"] = np.array([t.y_mm for t in tiles])
            rec["tile_z"] = np.array([t.z_level for t in tiles])
            rec["tile_channel"] = np.array([t.channel for t in tiles])
            rec["canvas"] = canvas
            for c, ff in (flatfields or {}).items():
                rec[f"flat_{c}"] = ff
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **rec)
        print(f"{name}: h={tuple(s.h_shift)} v={tuple(s.v_shift)} truth_h={truth['h_shift']} "
              f"truth_v={truth['v_shift']} canvas={canvas.shape}")


def pcc_known_answers():
    """Strip-shaped phase-correlation cases through the reference's own calculate_*_shift."""
    import contextlib
    import io
    ref_proc, ref_params = ref_shim.import_reference()
    from multiprocessing import Event, Queue
    rec = {}
    with tempfile.TemporaryDirectory() as tmp:
        p = ref_params.StitchingParameters(input_folder=tmp, use_registration=True)
        s = ref_proc.StitcherProcess(p, Queue(), Queue(), Queue(), Event())
        s.output_folder = tmp
        rng = np.random.default_rng(99)
        cases = []
        for i, (h, w, ov, dy, dx) in enumerate([(256, 320, 34, 2, -3), (300, 200, 22, -1, 2), (128, 128, 13, 0, 0),
                                                (214, 107, 11, 3, 1)]):
            world = synth.make_world(h * 2 + 64, w * 2 + 64, rng)
            a = np.clip(world[20:20 + h, 20:20 + w] + rng.normal(0, 30, (h, w)), 0, 65535).astype(np.uint16)
            bh = np.clip(world[20 + dy:20 + dy + h, 20 + w - ov + dx:20 + 2 * w - ov + dx] + rng.normal(0, 30, (h, w)),
                         0, 65535).astype(np.uint16)
            bv = np.clip(world[20 + h - ov + dy:20 + 2 * h - ov + dy, 20 + dx:20 + dx + w] + rng.normal(0, 30, (h, w)),
                         0, 65535).astype(np.uint16)
            with contextlib.redirect_stdout(io.StringIO()):
                hs = s.calculate_horizontal_shift(a, bh, ov)
                vs = s.calculate_vertical_shift(a, bv, ov)
            rec[f"a_{i}"], rec[f"bh_{i}"], rec[f"bv_{i}"] = a, bh, bv
            rec[f"ov_{i}"] = np.array(ov)
            rec[f"h_{i}"] = np.array(hs, dtype=np.int64)
            rec[f"v_{i}"] = np.array(vs, dtype=np.int64)
            cases.append(i)
            print(f"pcc case {i}: {h}x{w} ov={ov} true=({dy},{dx}) -> h={hs} v={vs}")
        rec["n"] = np.array(len(cases))
    np.savez_compressed(os.path.join(OUT, "shift_calls.npz"), **rec)


if __name__ == "__main__":
    if not ref_shim.reference_available():
        sys.exit("reference not present; goldens can only be regenerated in the build container")
    for name, kw in CASES.items():
        run_case(name, kw)
    for name, kw in FULL.items():
        run_case(name, kw, store_arrays=False)
    pcc_known_answers()
