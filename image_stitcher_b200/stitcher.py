"""``Stitcher`` -- the reference's second orchestrator (``stitcher.Stitcher(QThread)``, stitcher.py:31-1299) over the
same CUDA hot path.  The reference keeps AST-identical copies of the arithmetic methods in ``stitcher.py`` and
``stitcher_process.py``; here the arithmetic lives once, in :class:`StitcherProcess`, and this
class only swaps the plumbing: Qt-style signals instead of queues, ``params.validate()`` in the constructor
(stitcher.py:44), 512 x 512 chunks (stitcher.py:235), and a synchronous ``run()`` (``stitcher_cli.py:112`` calls it
directly).  With qtpy installed the signals are real ``Signal`` objects on a ``QThread``; without it they are small
callback lists with the same ``connect`` / ``emit`` surface.
"""
from __future__ import annotations

import os
import time

from .stitcher_parameters import StitchingParameters
from .stitcher_process import StitcherProcess


class _Signal:
    """``connect(fn)`` / ``emit(*args)`` stand-in for ``qtpy.QtCore.Signal`` when Qt is not installed."""

    def __init__(self, *types):
        self._slots = []

    def connect(self, fn):
        self._slots.append(fn)

    def emit(self, *args):
        for fn in list(self._slots):
            fn(*args)


class Stitcher(StitcherProcess):
    def __init__(self, params: StitchingParameters):
        params.validate()                                     # stitcher.py:44
        super().__init__(params, None, None, None, None)
        self.update_progress = _Signal(int, int)              # stitcher.py:33-37
        self.getting_flatfields = _Signal()
        self.starting_stitching = _Signal()
        self.starting_saving = _Signal(bool)
        self.finished_saving = _Signal(str, object)
        self.chunks = (1, 1, 1, 512, 512)                     # stitcher.py:235

    # the queue protocol of StitcherProcess mapped onto the signals
    def emit_progress(self, current: int, total: int):
        self.update_progress.emit(current, total)

    def emit_status(self, status: str, is_saving: bool = False):
        print(f"STATUS: {status}")

    def emit_complete(self, output_path: str, dtype):
        self.finished_saving.emit(output_path, dtype)

    def check_stop(self):                                     # a QThread is stopped from outside; nothing to poll
        return

    def start(self):                                          # QThread.start(): here simply run in the calling thread
        self.run()

    def run(self):
        """Same sequence as stitcher.py:1226-1299."""
        stime = time.time()
        try:
            self.get_timepoints()
            self.extract_acquisition_parameters()
            self.get_pixel_size()
            self.parse_acquisition_metadata()
            os.makedirs(self.output_folder, exist_ok=True)
            if self.apply_flatfield and not self.flatfields:
                self.getting_flatfields.emit()
                self.get_flatfields()
            if self.use_registration:
                self.calculate_shifts(self.timepoints[0], self.regions[0])
            final_path = ""
            for timepoint in self.timepoints:
                os.makedirs(os.path.join(self.output_folder, f"{timepoint}_stitched"), exist_ok=True)
                for region in self.regions:
                    self.starting_stitching.emit()
                    stitched = self.stitch_region(timepoint, region)
                    if not self.output_format.endswith(".zarr"):
                        raise RuntimeError("OME-TIFF output relies on the reference's third-party writers "
                                           "(out of scope, SURVEY.md section 2); use .ome.zarr")
                    self.starting_saving.emit(False)
                    final_path = self.save_region_ome_zarr(timepoint, region, stitched)
            self.finished_saving.emit(final_path, self.dtype)
            print(f"Processing complete. Total time: {time.time() - stime:.1f}s")
        finally:
            if self._ctx is not None:
                self._ctx.close()
                self._ctx = None
