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
  restates the published algorithms.  Since round 2 it is pinned independently of those
  algorithms: ``tests/golden/make_geodesy_mp.py`` evaluates the maps from their mathematical
  definitions in mpmath at 50 digits (transverse Mercator as the analytic continuation of the
  meridian arc, Newton iteration for ECEF -> geodetic) and ``tests/test_geodesy_pin.py`` holds this
  module to it within 5e-9 m over the five benchmark AOIs (1e-8 m world-wide) -- float64 noise.
  What stays unpinned is only what no reference RUN exists for: that the installed pymap3d / PROJ
  builds themselves agree with the exact maps to the same level (their documented accuracy).
"""
