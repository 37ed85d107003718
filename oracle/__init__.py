"""CPU oracle for the aggregate_2p5d hot path.  TEST INFRASTRUCTURE ONLY.

This package restates, in numpy (+ the real ``cv2.medianBlur``), the algorithm of the
reference's depth-map -> fused-DSM path.  It is the checker for the CUDA path, never the
product: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it.  Nothing under ``vissatsatellitestereo_b200/`` imports it.

Parity status
-------------
* The reference's OWN code on this path (unprojection block, ``proj_to_grid``, fusion block,
  ``read_array``) is pinned: ``tests/golden/make_golden.py`` imports and executes those reference
  functions from ``/root/reference`` (with stand-ins only for absent third-party imports) and
  freezes their outputs in ``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` holds the
  oracle to them bit-for-bit.
* The third-party arithmetic the reference delegates to (pymap3d 1.7.15 ``enu2geodetic`` /
  ``geodetic2enu``, PROJ 6.2 ``etmerc`` through pyproj 2.4.0, ``utm`` 0.4.2 zone rule,
  numpy_groupies 0.9.9 ``nanmax``) is NOT installable here (no network).  ``oracle/geodesy.py``
  restates the published algorithms; it is anchored on known answers (PROJ's documented
  ``echo 12 56 | proj +proj=utm +zone=32`` -> 687071.44 6210141.33, pymap3d's test triple
  (42,-82,200) -> ECEF, meridian-arc quadrature, round trips) but has no reference-run
  golden vectors: for that slice **parity is unpinned**.
"""
