"""CPU oracle for the PETITE shower-stepping hot path.  TEST INFRASTRUCTURE ONLY.

This package is a plain NumPy/SciPy restatement of the reference algorithm
(kjkellyphys/PETITE, files cited function by function).  It exists to CHECK the
CUDA engine in ``petite_b200``; nothing in the product path may import it.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs are allowed to call into ``oracle``.

Pinning status
--------------
* Everything the reference implements itself (integrands, form factors,
  kinematics, multiple scattering, Particle bookkeeping, the Shower /
  DarkShower drivers, table construction) is pinned against golden vectors
  produced by importing the UNMODIFIED reference modules in the build
  container (``tests/golden/make_golden.py``; the vectors are committed under
  ``tests/golden/``).
* The VEGAS adaptive-map transform lives in the third-party ``vegas`` package
  (``vegas>=5.4.2``, reference ``setup.cfg:30``), which is not installed and
  not vendored: **parity at the vegas boundary is unpinned**.  The restatement
  in ``oracle/vegasmap.py`` follows the published algorithm (Lepage, J. Comput.
  Phys. 27 (1978) 192 and the vegas 5.x documentation): piecewise-linear map
  y -> x on a stored node grid with jacobian ninc * dx.  The only weak pin is
  statistical: sum(wgt * f) through the shipped maps reproduces the shipped
  ``sm_xsec.pkl`` / ``dark_xsec.pkl`` rows (tests/test_oracle_xsec.py).
"""
