#!/usr/bin/env python
"""REFERENCE-side ensembles for the statistical parity tests: the UNMODIFIED reference classes, stream mode.

    python tests/golden/make_ensemble_ref.py [config ...] [--workers N]   # -> tests/golden/ensemble_ref.npz

Build container only (needs /root/reference).  Every shower is one call of the reference's own
``Shower.generate_shower`` (src/PETITE/shower.py:603-708) - and, for the dark configurations, of
``DarkShower.generate_dark_shower`` (src/PETITE/dark_shower.py:806-849) on that shower - imported through
``_refstub`` (stub ``vegas``: the one-hypercube sampler of oracle/vegasmap.py drawing from NumPy's global stream, the
only code on the path that is not the reference's).  Shower ``i`` of a configuration seeds both global generators the
reference draws from (``numpy.random`` and Python's ``random``, SURVEY Q-9) with ``seed0 + i``.

Per-shower summaries only (definitions: tests/ensemble_stats.py, the same ones the GPU side computes from its stack).
``ensemble.npz`` (tests/golden/make_ensemble.py) holds the same observables from the oracle in COUNTER mode;
tests/test_oracle_golden.py compares the two (counter-vs-stream KS test) and tests/test_gpu_ensemble.py compares the GPU
with both.
"""
import os
import random
import sys
import time
import warnings
from multiprocessing import Pool

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
from tests.ensemble_stats import REF_CONFIGS, SPEC_EDGES, config_primaries   # noqa: E402

REFDIR = "/tmp/petite_refdata/"
REFDIR_400 = "/tmp/petite_refdata_400GeV/"
SEED0 = 7_000_000
_STATE = {}
_PRIM = {}


def _summarise_sm(plist):
    E = np.array([float(q.get_p0()[0]) for q in plist]); pid = np.array([q.get_ids()["PID"] for q in plist])
    p0 = np.array([np.asarray(q.get_p0(), dtype=float) for q in plist]); rf = np.array([np.asarray(q.get_rf(), dtype=float) for q in plist])
    sec = np.arange(len(plist)) > 0
    g = pid == 22
    el = (np.abs(pid) == 11) & sec
    th = np.arctan2(np.hypot(p0[:, 1], p0[:, 2]), p0[:, 3])
    return dict(mult=len(plist), n_gamma=int(g.sum()), n_eplus=int((pid == -11).sum()), E_gamma=float(E[g & sec].sum()),
                Emax_sec=float(E[sec].max()) if sec.any() else 0.0,
                z_mean=float(np.sum(E * rf[:, 2]) / np.sum(E)), rT_mean=float(np.mean(np.hypot(rf[:, 0], rf[:, 1]))),
                theta_e=float(th[el].mean()) if el.any() else 0.0,
                spec=np.histogram(E[g & sec], bins=SPEC_EDGES)[0].astype(np.float64))


def _summarise_dark(vs):
    if not vs:
        return dict(n_V=0, dyield=0.0, lw_med=-300.0, EV_mean=0.0, EV_max=0.0)
    w = np.array([float(v.get_ids()["weight"]) for v in vs]); E = np.array([float(v.get_p0()[0]) for v in vs])
    return dict(n_V=len(vs), dyield=float(w.sum()), lw_med=float(np.median(np.log10(np.maximum(w, 1e-300)))),
                EV_mean=float(E.mean()), EV_max=float(E.max()))


def _engine(name):
    if name not in _STATE:
        import _refstub
        _refstub.import_reference()
        warnings.filterwarnings("ignore")
        cfg = REF_CONFIGS[name]
        refdir = REFDIR_400 if cfg.get("data") == "data_400GeV" else REFDIR
        if cfg.get("mV") is None:
            from PETITE.shower import Shower
            _STATE[name] = Shower(refdir, cfg["material"], cfg["E_min"])
        else:
            from PETITE.dark_shower import DarkShower
            _STATE[name] = DarkShower(refdir, cfg["material"], cfg["E_min"], cfg["mV"],
                                      active_processes=cfg.get("active"))
    return _STATE[name]


def _one(args):
    name, i = args
    from PETITE.particle import Particle    # after _engine() installed the stubs in this worker
    cfg = REF_CONFIGS[name]
    s = _engine(name)
    if name not in _PRIM:
        _PRIM[name] = config_primaries(cfg, cfg["n_ref"])
    p, pid, mass = _PRIM[name]
    np.random.seed(SEED0 + i)
    random.seed(SEED0 + i)
    t0 = time.time()
    sm = s.generate_shower(Particle([float(v) for v in p[i]], [0.0, 0.0, 0.0], {"PID": int(pid[i]), "ID": 1, "mass": float(mass[i])}))
    row = _summarise_sm(sm)
    if cfg.get("mV") is not None:
        _, vs = s.generate_dark_shower(ExDir=list(sm))
        row.update(_summarise_dark(vs))
    row["seconds"] = time.time() - t0
    return row


def _init(name):
    _engine(name)


def main():
    args = [a for a in sys.argv[1:]]
    workers = os.cpu_count()
    if "--workers" in args:
        k = args.index("--workers")
        workers = int(args[k + 1])
        del args[k:k + 2]
    path = os.path.join(HERE, "ensemble_ref.npz")
    only = [a for a in args if a in REF_CONFIGS]
    out = dict(np.load(path)) if os.path.exists(path) else {}
    sys.path.insert(0, HERE)
    import make_golden            # builds the reference-format dict_dir from data/*.npz (sm_maps.pkl, dark_maps.pkl)
    make_golden.build_reference_dict_dir()
    if any(REF_CONFIGS[n].get("data") == "data_400GeV" for n in (only or REF_CONFIGS)):
        make_golden.build_reference_dict_dir(data=os.path.join(ROOT, "data_400GeV", ""), refdir=REFDIR_400, ref_sub="data_400GeV", materials=["lead"])
    for name, cfg in REF_CONFIGS.items():
        if only and name not in only:
            continue
        t0 = time.time()
        _engine(name)             # built once here (warms the reference's own weight / dRate caches in REFDIR); workers fork it
        with Pool(workers, initializer=_init, initargs=(name,)) as pool:
            rows = pool.map(_one, [(name, i) for i in range(cfg["n_ref"])], chunksize=1)
        for k in list(out):
            if k.startswith(name + "/"):
                del out[k]
        for k in rows[0]:
            out[f"{name}/{k}"] = np.array([r[k] for r in rows])
        print(name, len(rows), f"{time.time() - t0:.0f} s wall", {k: float(np.mean(out[f'{name}/{k}'])) for k in rows[0] if np.ndim(rows[0][k]) == 0},
              flush=True)
        np.savez_compressed(path, **out)


if __name__ == "__main__":
    main()
