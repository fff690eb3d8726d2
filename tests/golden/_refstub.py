"""Import the UNMODIFIED reference (read-only /root/reference) in the build container.

Only used by the golden-vector generator scripts in this directory; never by a
test at run time (the GPU box has no /root/reference).  The reference cannot be
imported as-is because ``vegas`` and ``matplotlib`` are not installed, so two stub
modules are registered first:

* ``vegas``: ``lbatchintegrand`` = identity decorator; ``AdaptiveMap`` = node-grid
  holder (used when unpickling the shipped maps); ``Integrator`` = the oracle's
  restatement of the one-hypercube sampler (oracle/vegasmap.py), drawing its y
  points from NumPy's legacy global stream.  This is the UNPINNED boundary (see
  oracle/__init__.py) - everything else that runs is the reference's own code.
* ``matplotlib`` / ``matplotlib.pyplot``: empty modules (only imported, never used
  on the shower path).
"""
import os
import sys
import types

import numpy as np

REF_ROOT = os.environ.get("PETITE_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))

from oracle.vegasmap import AdaptiveMapStub, map_points  # noqa: E402


class _StubIntegrator:
    """One-hypercube VEGAS sampler with B = neval points per sweep (oracle/vegasmap.py)."""

    def __init__(self, map=None, **kw):
        self.map = map
        self.neval = int(kw.get("neval", 1000))
        self.kw = kw

    def set(self, **kw):
        if "neval" in kw:
            self.neval = int(kw["neval"])
        self.kw.update(kw)

    def _sweep(self):
        dim = self.map.dim
        y = np.random.random((self.neval, dim))
        x, jac = map_points(self.map.grid, y)
        return x, jac / self.neval

    def random(self):
        x, wgt = self._sweep()
        for i in range(len(wgt)):
            yield x[i], wgt[i]

    def random_batch(self):
        yield self._sweep()


def install():
    if "vegas" not in sys.modules:
        vg = types.ModuleType("vegas")
        vg.lbatchintegrand = lambda cls: cls
        vg.batchintegrand = lambda cls: cls
        vg.AdaptiveMap = AdaptiveMapStub
        vg.Integrator = _StubIntegrator
        sub = types.ModuleType("vegas._vegas")
        sub.AdaptiveMap = AdaptiveMapStub
        vg._vegas = sub
        sys.modules["vegas"] = vg
        sys.modules["vegas._vegas"] = sub
    if "matplotlib" not in sys.modules:
        mpl = types.ModuleType("matplotlib")
        plt = types.ModuleType("matplotlib.pyplot")
        mpl.pyplot = plt
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.pyplot"] = plt
    src = os.path.join(REF_ROOT, "src")
    if src not in sys.path:
        sys.path.insert(0, src)


def import_reference():
    install()
    import PETITE  # noqa: F401
    return PETITE
