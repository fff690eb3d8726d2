"""VEGAS adaptive-map restatement (oracle).  TEST INFRASTRUCTURE - see oracle/__init__.py.

PARITY UNPINNED at this boundary: the arithmetic lives in third-party ``vegas``
(>=5.4.2, reference setup.cfg:30), absent from /root/reference.  What is restated
here is the published piecewise-linear map (Lepage 1978; vegas 5.x docs,
"AdaptiveMap"):  for y in [0,1)^dim and a stored node grid g[d][0..ninc_d],

    i   = floor(y_d * ninc_d)          (clamped to ninc_d-1)
    x_d = g[d][i] + (g[d][i+1]-g[d][i]) * (y_d*ninc_d - i)
    jac = prod_d ninc_d * (g[d][i+1]-g[d][i])

and the reference's call sites (shower.py:433,453-457; dark_shower.py:669,690-694;
utilities/find_maxes.py:74-76,105-110): with ``max_nhcube=1`` / ``nstrat=1`` there is
one hypercube, each sweep yields ``B`` points with weight ``wgt = jac / B``.
``B`` (points per sweep) is not recoverable from the reference tree; this project
fixes ``B = neval`` (= 300 in every shipped table) for BOTH the max_F construction
and the sampler, which makes the sampled distribution independent of ``B``.

The shipped ``<Proc>_AdaptiveMaps.npy`` files are pickles of
``vegas._vegas.AdaptiveMap`` whose reduce-args are one list-of-lists of node
positions; they are read with a stub class.
"""
import pickle

import numpy as np
import numpy.lib.format as _fmt


class AdaptiveMapStub:
    """Stands in for vegas._vegas.AdaptiveMap when unpickling (reduce args = node grid)."""

    def __init__(self, grid, *a, **k):
        self.grid = [np.asarray(g, dtype=np.float64) for g in grid]

    @property
    def dim(self):
        return len(self.grid)

    def __reduce__(self):
        return (AdaptiveMapStub, ([g.tolist() for g in self.grid],))


class _Unpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if module.startswith("vegas"):
            return AdaptiveMapStub
        return super().find_class(module, name)


def load_pickle(path):
    with open(path, "rb") as f:
        return _Unpickler(f).load()


def load_adaptive_maps_npy(path):
    """Read a reference ``<Proc>_AdaptiveMaps.npy`` -> list of (params_dict, [grid_d arrays])."""
    with open(path, "rb") as f:
        ver = _fmt.read_magic(f)
        if ver == (1, 0):
            _fmt.read_array_header_1_0(f)
        else:
            _fmt.read_array_header_2_0(f)
        arr = _Unpickler(f).load()
    return [(dict(row[0]), row[1].grid) for row in arr]


def map_points(grid, y):
    """y: (..., dim) in [0,1) -> (x (..., dim), jac (...))."""
    y = np.asarray(y, dtype=np.float64)
    x = np.empty_like(y)
    jac = np.ones(y.shape[:-1], dtype=np.float64)
    for d, g in enumerate(grid):
        ninc = len(g) - 1
        yn = y[..., d] * ninc
        i = np.minimum(np.floor(yn).astype(np.int64), ninc - 1)
        inc = g[i + 1] - g[i]
        x[..., d] = g[i] + inc * (yn - i)
        jac = jac * (inc * ninc)
    return x, jac
