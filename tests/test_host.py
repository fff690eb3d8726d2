"""Host-side logic (no GPU): table readers, interpolant, Particle, the C-ABI library's symbols."""
import ctypes
import os
import re

import numpy as np
import pytest

from tests.conftest import DATA, ROOT


def test_library_exports_every_declared_symbol(built_library):
    lib = ctypes.CDLL(built_library)
    header = open(os.path.join(ROOT, "include", "petite_b200.h")).read()
    declared = set(re.findall(r"\b(pb_[a-z_0-9]+)\s*\(", header))
    assert {"pb_create", "pb_destroy", "pb_upload_nsigma", "pb_upload_maps", "pb_run_showers", "pb_probe"} <= declared
    for name in declared:
        assert hasattr(lib, name), name
    lib.pb_version.restype = ctypes.c_char_p
    assert b"sm_100a" in lib.pb_version()


def test_capi_struct_layouts(built_library):
    from petite_b200 import _capi
    assert set(_capi.SIGNATURES) >= {"pb_run_showers", "pb_probe"}
    assert ctypes.sizeof(_capi.pb_stack) == 7 * 8            # p0, r0w, pf, rf, ids (packed meta + key + weight), aux, capacity
    assert ctypes.sizeof(_capi.pb_primaries) == 8 * 8
    assert ctypes.sizeof(_capi.pb_profile) == 8 * (8 + 8 + 16 + 16)
    assert ctypes.sizeof(_capi.pb_counters) == 10 * 8
    assert ctypes.sizeof(_capi.pb_config) == 8 * (10 + 5 + 1 + 6 + 1)
    from petite_b200 import dark_setup
    assert dark_setup.CALL_DTYPE.itemsize == 48                # struct pb_quad_call: 6 x int32, 3 x double


def test_engine_fails_loudly_without_gpu(built_library):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from petite_b200.shower import Shower
    with pytest.raises(RuntimeError):
        Shower(DATA, "graphite", 0.010)


def test_tables_and_mapset():
    from petite_b200 import tables as tb
    maps = tb.load_sm_maps(DATA, "lead")
    assert set(maps) == set(tb.SM_PROCESSES)
    b = maps["Brem"]
    assert b.dim == 4 and list(b.ninc) == [960, 1000, 1000, 1000] and b.grid.shape == (100, 3964)
    assert b.neval == 300 and b.Eg_min == 0.001 and b.Ee_min == 0.005
    assert np.all(np.isfinite(b.max_F)) and np.all(b.max_F >= 0)
    nodes = b.axis_nodes(10, 1)
    assert len(nodes) == 1001 and nodes[0] == 0.0 and abs(nodes[-1] - 2.0) < 1e-12 and np.all(np.diff(nodes) > 0)
    xs = tb.load_sm_xsec(DATA, "graphite")
    assert np.allclose(xs["Brem"][0], [0.0016, 11.883294458724587]) and np.allclose(xs["Brem"][-1], [100.0, 14319.795619370949])
    with pytest.raises(Exception, match="Target Material is not in library"):
        tb.load_sm_xsec(DATA, "unobtainium")
    assert tb.list_dark_masses(DATA) == [0.003, 0.01, 0.03, 0.1, 0.3, 1.0]


def test_linear_table_matches_scipy():
    from scipy.interpolate import interp1d
    from petite_b200.shower import LinearTable
    rng = np.random.default_rng(0)
    x = np.sort(rng.random(50)) * 10
    y = rng.random(50)
    q = np.concatenate([rng.random(200) * 12 - 1, x[:5], [x[0], x[-1]]])
    want = interp1d(x, y, fill_value=0.0, bounds_error=False)(q)
    assert np.array_equal(LinearTable(x, y)(q), want)


def test_particle_api_matches_reference_conventions():
    from petite_b200 import Particle
    from petite_b200.constants import m_electron
    p = Particle([10.0, 0, 0, np.sqrt(100 - m_electron ** 2)], [0, 0, 0], {"PID": 11, "ID": 0})
    ids = p.get_ids()
    assert ids["mass"] == 0.000511 and ids["generation_process"] == "Input" and ids["stability"] == "stable"   # Q-7
    assert p.get_ended() is False and np.array_equal(p.get_pf(), p.get_p0())
    q = Particle(5.0, id_dictionary={"PID": 13})
    assert abs(q.get_p0()[3] ** 2 - (25 - 0.1056583755 ** 2)) < 1e-12
    p.lose_energy(20.0)
    assert list(p.get_pf()) == [0.000511, 0.0, 0.0, 0.0]
    with pytest.raises(ValueError):
        p.set_ended("yes")


def test_reference_order_reconstruction():
    """ShowerBatch.reference_order on a hand-built stack: waves unordered inside, two showers interleaved."""
    from petite_b200.shower import ShowerBatch

    class T:
        def __init__(self, a):
            self.a = a

        def __getitem__(self, s):
            return self

        def cpu(self):
            return self

        def numpy(self):
            return self.a
    # slots: 0,1 primaries (showers 0,1); wave 1: slots 2..5; wave 2: slots 6..8
    parent = np.array([-1, -1, 1, 0, 0, 1, 3, 2, 3])
    bit = np.array([0, 0, 1, 1, 0, 0, 1, 0, 0])
    gen = np.array([0, 0, 1, 1, 1, 1, 2, 2, 2])
    shower = np.array([0, 1, 1, 0, 0, 1, 0, 1, 0])
    n = len(parent)
    meta = np.stack([np.full(n, 11), parent, (gen << 16) | (bit << 15), shower], axis=1).astype(np.int32)
    z4 = np.zeros((n, 4))
    t = {"p0": T(z4), "r0w": T(z4), "pf": T(z4), "rf": T(z4), "key": T(np.zeros((n, 2), np.int32)), "meta": T(meta),
         "aux": T(np.zeros((n, 2), np.int32))}
    b = ShowerBatch(None, t, n, {}, 2, 0)
    order, offs = b.reference_order()
    assert list(offs) == [0, 5, 9]
    assert list(order[:5]) == [0, 4, 3, 8, 6]      # shower 0: primary; daughters bit0, bit1; then slot 3's daughters
    assert list(order[5:]) == [1, 5, 2, 7]


def test_reference_format_dict_dir_is_accepted(tmp_path):
    """Drop-in: a PETITE-style dict_dir (sm_xsec.pkl, sm_maps.pkl with pickled vegas AdaptiveMap objects, dark_*.pkl) must
    load without vegas and give the same tables as the repacked .npz files."""
    import pickle
    import sys
    import types
    from petite_b200 import tables as tb

    class AdaptiveMap:                       # stands in for vegas._vegas.AdaptiveMap when WRITING the fixture
        def __init__(self, grid):
            self.grid = grid

        def __reduce__(self):
            return (AdaptiveMap, ([list(map(float, g)) for g in self.grid],))
    AdaptiveMap.__module__, AdaptiveMap.__qualname__ = "vegas._vegas", "AdaptiveMap"
    mod = types.ModuleType("vegas._vegas"); mod.AdaptiveMap = AdaptiveMap
    pkg = types.ModuleType("vegas"); pkg._vegas = mod
    saved = {k: sys.modules.get(k) for k in ("vegas", "vegas._vegas")}
    sys.modules["vegas"], sys.modules["vegas._vegas"] = pkg, mod
    try:
        npz = tb.load_sm_maps(DATA, "lead")
        xs = tb.load_sm_xsec(DATA, "lead")
        d = str(tmp_path) + "/"
        maps = {P: [[float(ms.E[i]), {"neval": ms.neval, "max_F": {"lead": float(ms.max_F[i])}, "Eg_min": ms.Eg_min, "Ee_min": ms.Ee_min,
                                      "adaptive_map": AdaptiveMap([ms.axis_nodes(i, k) for k in range(ms.dim)])}] for i in range(0, 100, 33)]
                for P, ms in npz.items()}
        pickle.dump(maps, open(d + "sm_maps.pkl", "wb"))
        pickle.dump({P: {"lead": [list(r) for r in xs[P]]} for P in xs}, open(d + "sm_xsec.pkl", "wb"))
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    assert "vegas" not in sys.modules                       # reading must not need it
    got = tb.load_sm_maps(d, "lead")
    for P, ms in got.items():
        assert ms.dim == npz[P].dim and list(ms.ninc) == list(npz[P].ninc)
        assert np.array_equal(ms.grid, npz[P].grid[::33]) and np.array_equal(ms.max_F, npz[P].max_F[::33])
        assert ms.neval == 300
    gx = tb.load_sm_xsec(d, "lead")
    assert all(np.array_equal(gx[P], xs[P]) for P in xs)
    with pytest.raises(Exception, match="Target Material is not in library"):
        tb.load_sm_xsec(d, "graphite")
