"""Run the UNMODIFIED reference ``stitcher_process.py`` under import shims.

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.  Works only where
``/root/reference`` exists (the build container); it is used to *generate*
``tests/golden/*.npz`` (``tests/golden/make_golden.py``) and by the optional
``not gpu`` tests that re-check the goldens when the reference is present.
Nothing that runs on the GPU box imports this module.

The reference imports a dozen third-party packages at module top
(``stitcher_process.py:11-27``) that are not installed here.  Stub modules are
registered in ``sys.modules`` for them:

* ``skimage.registration.phase_cross_correlation`` -> ``oracle.pcc_ref`` restatement
* ``dask.array.zeros`` -> ``numpy.zeros`` (NumPy slice assignment reproduces dask's
  sequential last-writer-wins ``__setitem__``)
* ``dask_image.imread.imread`` -> ``cv2.imread(..., IMREAD_UNCHANGED)[None]`` (colour images flipped BGR -> RGB)
* writers / BaSiC / pyvips / zarr -> inert placeholders (never called by the hot path)
"""
from __future__ import annotations

import contextlib
import io
import os
import sys
import types
from multiprocessing import Event, Queue

import numpy as np

REFERENCE_DIR = os.environ.get("STITCH_REFERENCE_DIR", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_DIR, "stitcher_process.py"))


def _mod(name: str, **attrs) -> types.ModuleType:
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


class _Inert:
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        raise RuntimeError("inert shim object called: this code path is outside the oracle's scope")


def install_stubs() -> None:
    import cv2
    from . import pcc_ref

    def _pcc(reference_image, moving_image, upsample_factor=1, **kw):
        return pcc_ref.phase_cross_correlation(np.asarray(reference_image), np.asarray(moving_image),
                                               upsample_factor=upsample_factor)

    def _imread(path):
        img = cv2.imread(path, cv2.IMREAD_UNCHANGED)
        if img is None:
            raise FileNotFoundError(path)
        if img.ndim == 3 and img.shape[2] == 3:
            img = np.ascontiguousarray(img[:, :, ::-1])     # OpenCV decodes BGR; dask_image (pims) yields RGB
        return img[None]

    class _DaskArray:  # only used in isinstance() checks (stitcher_process.py:1993)
        pass

    def _zeros(shape, dtype=float, chunks=None):
        return np.zeros(shape, dtype=dtype)

    sk = _mod("skimage")
    sk.registration = _mod("skimage.registration", phase_cross_correlation=_pcc)
    sk.exposure = _mod("skimage.exposure")
    dk = _mod("dask")
    dk.array = _mod("dask.array", Array=_DaskArray, zeros=_zeros)
    di = _mod("dask_image")
    di.imread = _mod("dask_image.imread", imread=_imread)
    for name in ("ome_zarr", "zarr", "imageio", "pyvips"):
        _mod(name)
    _mod("basicpy", BaSiC=_Inert)
    aics = _mod("aicsimageio", types=_mod("aicsimageio.types"))
    aics.writers = _mod("aicsimageio.writers", OmeTiffWriter=_Inert, OmeZarrWriter=_Inert)
    bio = _mod("bioio")
    bio.writers = _mod("bioio.writers", OmeTiffWriter=_Inert, OmeZarrWriter=_Inert)
    bio.writers.ome_zarr_writer_2 = _mod("bioio.writers.ome_zarr_writer_2", OmeZarrWriter=_Inert,
                                         compute_level_shapes=_Inert(), compute_level_chunk_sizes_zslice=_Inert())
    _mod("bioio_base", types=_mod("bioio_base.types"))


def import_reference():
    """Import the reference's ``stitcher_process`` + ``stitcher_parameters`` modules (read-only)."""
    if not reference_available():
        raise FileNotFoundError(f"reference not present at {REFERENCE_DIR}")
    install_stubs()
    if REFERENCE_DIR not in sys.path:
        sys.path.insert(0, REFERENCE_DIR)
    sys.dont_write_bytecode = True     # never write __pycache__ into the read-only reference
    import stitcher_parameters as ref_params   # noqa: E402
    import stitcher_process as ref_proc        # noqa: E402
    return ref_proc, ref_params


def make_reference_stitcher(input_folder: str, output_folder: str, *, quiet: bool = True, **params):
    """Construct a reference ``StitcherProcess`` and run its metadata phase (1964-1969)."""
    ref_proc, ref_params = import_reference()
    p = ref_params.StitchingParameters(input_folder=input_folder, **params)
    s = ref_proc.StitcherProcess(p, Queue(), Queue(), Queue(), Event())
    s.output_folder = output_folder
    os.makedirs(output_folder, exist_ok=True)
    sink = io.StringIO() if quiet else sys.stdout
    with contextlib.redirect_stdout(sink):
        s.get_timepoints()
        s.extract_acquisition_parameters()
        s.get_pixel_size()
        s.parse_acquisition_metadata()
    return s


def run_reference(s, *, flatfields=None, quiet: bool = True):
    """``calculate_shifts`` (if registration is on) then ``stitch_region`` for every region."""
    sink = io.StringIO() if quiet else sys.stdout
    out = {}
    with contextlib.redirect_stdout(sink):
        if flatfields is not None:
            s.flatfields = dict(flatfields)
        if s.use_registration:
            s.calculate_shifts(s.timepoints[0], s.regions[0])
        for t in s.timepoints:
            for region in s.regions:
                out[(int(t), region)] = np.asarray(s.stitch_region(t, region))
    return out
