"""pytest configuration: the ``gpu`` marker and shared golden-fixture helpers."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    """Load a golden case -> (dict of arrays, RegionState, tiles in paste order)."""
    from oracle.stitch_ref import RegionState, TileRec
    g = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False))
    kw = eval(str(g["kwargs"]), {"__builtins__": {}}, {"dict": dict})  # repr() of a plain kwargs dict
    tiles = []
    if "tiles" in g:
        for i, nm in enumerate(g["tile_names"]):
            nm = str(nm)
            fov = int(nm.split("_")[1])
            tiles.append(TileRec(x_mm=float(g["tile_x_mm"][i]), y_mm=float(g["tile_y_mm"][i]),
                                 z_level=int(g["tile_z"][i]), channel=str(g["tile_channel"][i]),
                                 pixels=g["tiles"][i], fov=fov, name=nm))
    mono = [str(c) for c in g["monochrome_channels"]]
    flat = {int(k.split("_")[1]): g[k] for k in g if k.startswith("flat_")}
    st = RegionState(tile_h=kw["tile_h"], tile_w=kw["tile_w"], pixel_size_um=float(g["pixel_size_um"]),
                     pixel_binning=2, monochrome_channels=mono, channel_names=list(mono),
                     num_z=kw.get("num_z", 1), use_registration=kw.get("use_registration", False),
                     apply_flatfield=kw.get("apply_flatfield", False),
                     scan_pattern=kw.get("scan_pattern", "Unidirectional"),
                     registration_channel=kw.get("registration_channel", ""), flatfields=flat)
    if tiles:
        st.dtype = np.dtype(tiles[0].pixels.dtype)
    return g, st, tiles, kw


SMALL_GOLDENS = ["reg_2x2_mono", "reg_3x3_spattern_flat", "coord_3x4_flat64", "coord_2x2_plain", "reg_2x3_negdrift",
                 "reg_2x2_u8_flat"]


RGB_GOLDENS = ["coord_2x3_rgb_u8"]       # colour tiles: only through the paths that split planes (oracle, StitcherProcess)


@pytest.fixture(scope="session")
def golden_loader():
    return load_golden
