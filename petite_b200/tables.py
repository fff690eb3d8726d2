"""On-disk table readers for the shower engine (host side, NumPy only).

Two directory layouts are accepted by ``Shower(dict_dir, ...)``:

* the reference layout (``sm_xsec.pkl``, ``sm_maps.pkl``, ``dark_xsec.pkl``, ``dark_maps.pkl``,
  ``dark_weights.pkl``, ``dark_drate.pkl``; reference shower.py:153-175, dark_shower.py:148-217 and
  SURVEY.md 3.5).  ``*_maps.pkl`` hold pickled ``vegas._vegas.AdaptiveMap`` objects, which reduce to a
  list-of-lists of node positions and are read with a stub class, so ``vegas`` is not required;
* this project's pickle-free ``.npz`` layout written by ``tools/pack_reference_data.py`` plus the
  ``*_maxF.npz`` tables (the reference's ``*_maps.pkl`` are missing upstream, so max_F had to be
  regenerated).

Either way the result is a :class:`MapSet` per process: the flattened fp64 node arena that is uploaded
verbatim to HBM (layout in DESIGN.md section 3).
"""
import os
import pickle
from dataclasses import dataclass

import numpy as np
import numpy.lib.format as _fmt

SM_PROCESSES = ["Brem", "Ann", "PairProd", "Comp", "Moller", "Bhabha", "MuonE", "MuonBrem"]
DARK_PROCESSES = ["DarkBrem", "DarkAnn", "DarkComp", "DarkMuonBrem"]
PROC_DIM = {"Brem": 4, "Ann": 1, "PairProd": 4, "Comp": 1, "Moller": 1, "Bhabha": 1, "MuonE": 1,
            "MuonBrem": 4, "DarkBrem": 3, "DarkAnn": 1, "DarkComp": 1, "DarkMuonBrem": 3}


def mv_tag(mV):
    return repr(float(mV))


class _NodeGrid:
    """Placeholder for ``vegas._vegas.AdaptiveMap`` while unpickling reference files."""

    def __init__(self, grid, *args, **kwargs):
        self.grid = [np.asarray(g, dtype=np.float64) for g in grid]


class _Unpickler(pickle.Unpickler):
    """The reference's table pickles hold dicts / lists of NumPy arrays and ``vegas`` objects (replaced by ``_NodeGrid``); nothing
    else is allowed to be constructed while loading one."""
    _ALLOWED = ("numpy", "builtins", "collections", "copyreg", "_codecs")

    def find_class(self, module, name):
        if module.startswith("vegas"):
            return _NodeGrid
        if module.split(".")[0] in self._ALLOWED and not (module == "builtins" and name in ("eval", "exec", "compile", "open", "__import__", "getattr")):
            return super().find_class(module, name)
        raise pickle.UnpicklingError(f"table pickle refers to {module}.{name}: only NumPy / builtin containers and vegas maps are expected")


def load_reference_pickle(path):
    with open(path, "rb") as f:
        return _Unpickler(f).load()


def load_adaptive_maps_npy(path):
    """``<Proc>_AdaptiveMaps.npy`` -> list of (params dict, [node array per dimension])."""
    with open(path, "rb") as f:
        version = _fmt.read_magic(f)
        (_fmt.read_array_header_1_0 if version == (1, 0) else _fmt.read_array_header_2_0)(f)
        rows = _Unpickler(f).load()
    return [(dict(r[0]), r[1].grid) for r in rows]


@dataclass
class MapSet:
    """All trained maps of one process: what ``pb_upload_maps`` ships to the GPU."""
    process: str
    E: np.ndarray        # (nE,) training energies
    ninc: np.ndarray     # (dim,) increments per axis; axis d has ninc[d]+1 nodes
    grid: np.ndarray     # (nE, sum(ninc+1)) fp64 node positions, axes concatenated
    max_F: np.ndarray    # (nE,) for the chosen material (NaN if unknown)
    neval: int
    Eg_min: float
    Ee_min: float

    @property
    def dim(self):
        return len(self.ninc)

    def axis_nodes(self, ie, d):
        off = int(np.sum(self.ninc[:d] + 1))
        return self.grid[ie, off:off + int(self.ninc[d]) + 1]


def _mapset_from_reference(process, rows, material):
    E = np.array([float(r[0]) for r in rows])
    grids = [r[1]["adaptive_map"].grid for r in rows]
    ninc = np.array([len(g) - 1 for g in grids[0]], dtype=np.int32)
    grid = np.stack([np.concatenate(g) for g in grids])
    info = rows[0][1]
    max_F = np.array([float(r[1]["max_F"][material]) for r in rows])
    return MapSet(process, E, ninc, grid, max_F, int(info["neval"]),
                  float(info.get("Eg_min", 0.001)), float(info.get("Ee_min", 0.005)))


def _mapset_from_npz(process, z, maxF_z, maxF_key):
    meta = z[f"{process}/meta"]
    E = z[f"{process}/E"]
    if maxF_z is not None and maxF_key in maxF_z:
        max_F = np.asarray(maxF_z[maxF_key], dtype=np.float64)
    else:
        max_F = np.full(len(E), np.nan)
    return MapSet(process, E, z[f"{process}/ninc"].astype(np.int32), z[f"{process}/grid"], max_F,
                  int(meta[0]), float(meta[1]), float(meta[2]))


def load_sm_xsec(dict_dir, material):
    """-> {process: (n,2) array of [E, sigma GeV^-2]} ; raises like shower.py:164-175."""
    pkl = dict_dir + "sm_xsec.pkl"
    out = {}
    if os.path.exists(pkl):
        with open(pkl, "rb") as f:
            d = pickle.load(f)
        for P in SM_PROCESSES:
            if P not in d:
                raise Exception("Process String does not match library")
            if material not in d[P]:
                raise Exception("Target Material is not in library")
            out[P] = np.asarray(d[P][material], dtype=np.float64)
        return out
    z = np.load(dict_dir + "sm_xsec.npz")
    for P in SM_PROCESSES:
        if not any(k.startswith(P + "/") for k in z.files):
            raise Exception("Process String does not match library")
        if f"{P}/{material}" not in z.files:
            raise Exception("Target Material is not in library")
        out[P] = z[f"{P}/{material}"]
    return out


def load_sm_maps(dict_dir, material):
    """-> {process: MapSet} for the 8 SM processes (reference shower.py:153-162, 210-215)."""
    pkl = dict_dir + "sm_maps.pkl"
    if os.path.exists(pkl):
        d = load_reference_pickle(pkl)
        out = {}
        for P in SM_PROCESSES:
            if P not in d:
                print(P)
                raise Exception("Process String does not match library")
            out[P] = _mapset_from_reference(P, d[P], material)
        return out
    z = np.load(dict_dir + "sm_maps.npz")
    mf = np.load(dict_dir + "sm_maxF.npz") if os.path.exists(dict_dir + "sm_maxF.npz") else None
    return {P: _mapset_from_npz(P, z, mf, f"{P}/{material}") for P in SM_PROCESSES}


def list_dark_masses(dict_dir):
    """Keys of dark_maps.pkl (dark_shower.py:148-155), in file order."""
    pkl = dict_dir + "dark_maps.pkl"
    if os.path.exists(pkl):
        return list(load_reference_pickle(pkl).keys())
    z = np.load(dict_dir + "dark_xsec.npz")
    seen = []
    for k in z.files:
        m = float(k.split("/")[0])
        if m not in seen:
            seen.append(m)
    return seen


def load_dark_xsec(dict_dir, mV, material):
    pkl = dict_dir + "dark_xsec.pkl"
    out = {}
    if os.path.exists(pkl):
        with open(pkl, "rb") as f:
            d = pickle.load(f)[mV]
        for P in DARK_PROCESSES:
            if P not in d:
                raise Exception("Process String does not match library")
            if material not in d[P]:
                raise Exception("Target Material is not in library")
            out[P] = np.asarray(d[P][material], dtype=np.float64)
        return out
    z = np.load(dict_dir + "dark_xsec.npz")
    for P in DARK_PROCESSES:
        key = f"{mv_tag(mV)}/{P}/{material}"
        if key not in z.files:
            raise Exception("Target Material is not in library")
        out[P] = z[key]
    return out


def load_dark_maps(dict_dir, mV, material, processes):
    pkl = dict_dir + "dark_maps.pkl"
    if os.path.exists(pkl):
        d = load_reference_pickle(pkl)[mV]
        out = {}
        for P in processes:
            if P not in d:
                print(P)
                raise Exception("Process String does not match library")
            out[P] = _mapset_from_reference(P, d[P], material)
        return out
    path = dict_dir + f"dark_maps_mV{mv_tag(mV)}.npz"
    if not os.path.exists(path):
        raise FileNotFoundError(f"no dark maps packed for mV={mV} under {dict_dir}")
    z = np.load(path)
    mf = np.load(dict_dir + "dark_maxF.npz") if os.path.exists(dict_dir + "dark_maxF.npz") else None
    return {P: _mapset_from_npz(P, z, mf, f"{mv_tag(mV)}/{P}/{material}") for P in processes}


def load_dark_weights(dict_dir, mV, material):
    """-> dict name -> (n,2) or None if (mV, material) is not cached (dark_shower.py:401-424)."""
    pkl = dict_dir + "dark_weights.pkl"
    if os.path.exists(pkl):
        with open(pkl, "rb") as f:
            d = pickle.load(f)
        if mV in d and material in d[mV]:
            return {k: np.asarray(v, dtype=np.float64) for k, v in d[mV][material].items()}
        return None
    path = dict_dir + "dark_weights.npz"
    if not os.path.exists(path):
        return None
    z = np.load(path)
    pre = f"{mv_tag(mV)}/{material}/"
    out = {k[len(pre):]: z[k] for k in z.files if k.startswith(pre)}
    return out or None


def load_dark_drate(dict_dir, mV, material):
    """-> dict name -> (E (n,), table (n,10,2)) or None if not cached (dark_shower.py:535-562)."""
    pkl = dict_dir + "dark_drate.pkl"
    if os.path.exists(pkl):
        with open(pkl, "rb") as f:
            d = pickle.load(f)
        if mV in d and material in d[mV]:
            out = {}
            for name, tab in d[mV][material].items():
                keys = sorted(tab.keys(), key=float)      # k_dark_prepare binary-searches the energies: ascending order
                out[name] = (np.array([float(k) for k in keys]),
                             np.stack([np.asarray(tab[k], dtype=np.float64) for k in keys]))
            return out
        return None
    path = dict_dir + "dark_drate.npz"
    if not os.path.exists(path):
        return None
    z = np.load(path)
    pre = f"{mv_tag(mV)}/{material}/"
    names = sorted({k[len(pre):].split("/")[0] for k in z.files if k.startswith(pre)})
    return {n: (z[f"{pre}{n}/E"], z[f"{pre}{n}/table"]) for n in names} or None
