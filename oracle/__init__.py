"""CPU oracle for the pyradiosky per-component evaluation hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``pyradiosky_b200/`` imports this package; only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs do, and there only as the checker / the timed CPU baseline, never as the thing shipped.

It is a numpy/scipy restatement of the reference's algorithm (the reference is pure Python, so the
restatement is too).  Each function cites the reference ``file:line`` it follows, relative to the
upstream pyradiosky tree.

Parity status (see DESIGN.md "Oracle"):
  * rotation (``sct_np``), Stokes->coherency, spectra, unit conversion: PINNED against the
    reference's own ``spherical_coords_transforms.py`` imported by path and against the golden values
    of the reference's tests / tutorial (``tests/golden/*.json``, ``tests/golden/make_golden.py``).
  * position chain (``erfa_np``): the reference delegates RA/Dec -> Alt/Az to astropy -> pyerfa ->
    liberfa, none of which is present in this image, so the ERFA ``atciqz``/``atioq``/``apco``
    algorithms are restated from their published form.  Pinned indirectly by the reference's golden
    instrumental-coherency matrix (test_skymodel.py:1468-1506, rtol 1e-5) and the tutorial's horizon
    mask (tutorial.rst:1129-1132).  Absolute Alt/Az at the micro-arcsecond level against astropy:
    PARITY UNPINNED (no astropy/pyerfa/IERS on this box).
  * HEALPix pixel centres (``healpix_np``): standard published ring/nested formulas; astropy-healpix
    absent, PARITY UNPINNED against it; pinned by self-consistency invariants only.
"""
