"""CPU oracle for the registration + fusion hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and there only as the checker (or as
the CPU arm being timed), never behind the CUDA path.

Parity status: **parity unpinned by the reference's own tests** -- the
reference (sohamazing/image-stitcher) ships no tests, golden vectors or
fixtures.  The oracle is pinned instead by

* ``tests/golden/*.npz`` -- outputs of the *unmodified* reference file
  ``/root/reference/stitcher_process.py`` executed in the build container under
  import shims (``oracle/ref_shim.py``; generator ``tests/golden/make_golden.py``);
* the known-answer tests of scikit-image's ``phase_cross_correlation`` test
  suite (recalled; scikit-image is an un-vendored, un-pinned dependency of the
  reference -- ``install_requirements.sh:54,65-67``).
"""
