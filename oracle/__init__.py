"""CPU oracle for the bundle-adjustment hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it, and only as the
checker or as the timed CPU baseline.  The shipped path
(``multicam_calibration_b200``) never imports this package and raises when its
CUDA library is missing.

Parity pinning: the reference has no tests, golden vectors or known-answer
fixtures of its own.  This oracle is therefore pinned
against OUTPUTS OF THE REFERENCE ITSELF, produced in the build container by
importing the unmodified ``/root/reference/multicam_calibration/{geometry,
bundle_adjustment}.py`` (``tests/golden/make_golden.py``) and committed as
``tests/golden/*.npz``.  The solver (scipy ``least_squares``) and the OpenCV
routines used by ``triangulate`` are third-party, un-vendored and un-pinned
by the reference (``setup.cfg:14-23``); the fixtures pin them to
scipy 1.18.1 / OpenCV 4.13.0 as installed in this image.
"""
