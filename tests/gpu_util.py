"""Helpers for the -m gpu tests: engine construction and the pb_probe call (everything goes through the C ABI)."""
import ctypes as C

import numpy as np

from tests.conftest import DATA

_CACHE = {}


def shower(material="graphite", min_energy=0.010, seed=0, **kw):
    from petite_b200.shower import Shower
    key = (material, min_energy, seed, tuple(sorted(kw.items())))
    if key not in _CACHE:
        _CACHE[key] = Shower(DATA, material, min_energy, seed=seed, **kw)
    return _CACHE[key]


def probe(sh, what, process, inp, out_cols):
    from petite_b200 import _capi as capi
    inp = np.ascontiguousarray(inp, dtype=np.float64)
    out = np.zeros((len(inp), out_cols))
    capi.check(sh._engine, capi.lib.pb_probe(sh._engine, what, process, capi.dptr(inp), len(inp), inp.shape[1],
                                             capi.dptr(out), out_cols))
    return out


def primaries(pid, E, n, mass=None, stability="stable"):
    from petite_b200 import Particle
    from petite_b200.constants import MASS
    m = MASS[pid]
    ids = {"PID": pid, "ID": 1, "mass": m if mass is None else mass, "stability": stability}
    return [Particle([E, 0, 0, np.sqrt(E ** 2 - m ** 2)], [0, 0, 0], dict(ids)) for _ in range(n)]
