"""VEGAS adaptive-map training on the GPU.

The reference trains one ``vegas.Integrator`` per (process, incoming energy) on host cores
(utilities/generate_integrators.py:49-183 -> all_processes.py:1160-1226, ``multiprocessing.Pool`` over energies) and
saves its ``AdaptiveMap``.  Here all energies of a process are trained together: ``pb_train_accumulate`` pushes uniform
y points through the current node grids and returns, per axis increment, the sum of |jac f|^p and the hit count; the
classic VEGAS refinement (Lepage 1978: smooth, damp with ``alpha``, re-bin to equal content) is a few vectorised NumPy
lines per iteration.  p = 2 is Lepage's variance criterion.  The maps are used for ACCEPT/REJECT sampling, whose cost is
max / mean of jac f, so the default here is p = 8 (``TRAIN_POWER``): it flattens the peaks that set max_F and gives maps with
2.3-2.8x the accept rate of the shipped ones (Brem, PairProd, MuonBrem; p = 2 reaches 0.3-0.5x because the reference's adaptive
stratification is not reproduced - profiles/r02_final/exp_train_pow*.log); the integral through the map is unbiased for any p.  Training is not bit-reproducible: the fp64 atomics of the sums are
unordered (1e-15 relative), the integrands have kinematic cuts, and twenty refinements amplify that to visibly different node
positions (tools/check_train_determinism.py) - single rows scatter, sets of rows do not; table sets are therefore committed, not
regenerated, where tests depend on them (data_400GeV/).  ``vegas`` itself is not available here, so this is the published algorithm, not a bit-level
reproduction of the third-party package; the acceptance test is the one the shipped tables allow - the integral through
the trained maps reproduces the shipped ``sm_xsec`` rows (tests/test_gpu_train.py).

    python -m petite_b200.train --xsec-from /path/to/data_400GeV/ --out data_400GeV/ --processes Brem,PairProd
"""
import argparse
import ctypes as C
import os

import numpy as np

from . import _capi as capi
from . import constants as K
from . import tables as tb

CODE = {"Brem": 0, "Ann": 1, "PairProd": 2, "Comp": 3, "Moller": 4, "Bhabha": 5, "MuonE": 6, "MuonBrem": 7,
        "DarkBrem": 8, "DarkAnn": 9, "DarkComp": 10, "DarkMuonBrem": 11}
NINC = {4: [960, 1000, 1000, 1000], 3: [1000, 1000, 1000], 1: [1000]}     # increments of the shipped maps
TRAIN_POWER = 8.0                # training weight |jac f|^p (module docstring)
TRAIN_SCHEDULE = [(10, 2_000_000, 1.0), (10, 8_000_000, 0.5)]      # (iterations, points per iteration, alpha)


def integration_range(process, E, mV=0.0, Eg_min=0.001, Ee_min=0.005):
    """Map domains, all_processes.py:1102-1158 (incl. the dimensionful MuonBrem domain, SURVEY Q-5)."""
    me = K.m_electron
    if process in ("Brem", "PairProd"):
        return [[0, 1], [0, 2], [-2, 2], [0, 1]]
    if process == "MuonBrem":
        maxdel = np.sqrt(E / max(me, mV))
        return [[max(Eg_min, mV), E - me], [0.0, maxdel], [0.0, maxdel], [0.0, 2 * np.pi]]
    if process in ("DarkBrem", "DarkMuonBrem"):
        return [[max(0.0, mV / E), 1.0 - me / E], [-12.0, np.log10(2.0)], [-20.0, 0.0]]
    if process in ("Comp", "Ann", "DarkComp", "MuonE"):
        return [[-1.0, 1.0]]
    if process == "DarkAnn":
        return [[0.0, 1.0]] if 2.0 * me * (E + me) > mV ** 2 else [[0.0, 0.0]]
    d = 2.0 * Ee_min / (E - me)
    return [[-1.0 + d, 1.0 - d]]


def refine(grid, d, cnt, alpha=0.5):
    """One VEGAS refinement of a node array ``grid`` (ninc+1,) from per-increment training data ``d`` / ``cnt``."""
    ninc = len(grid) - 1
    avg = np.where(cnt > 0, d / np.maximum(cnt, 1), 0.0)
    if ninc > 1:
        sm = np.empty(ninc)
        sm[0] = abs(7 * avg[0] + avg[1]) / 8
        sm[-1] = abs(7 * avg[-1] + avg[-2]) / 8
        sm[1:-1] = np.abs(6 * avg[1:-1] + avg[:-2] + avg[2:]) / 8
        tot = sm.sum()
        avg = sm / tot + 1e-300 if tot > 0 else np.full(ninc, 1e-300)
        with np.errstate(all="ignore"):
            avg = np.where(avg < 1.0, (-(1 - avg) / np.log(avg)) ** alpha, 1.0)
    cum = np.concatenate([[0.0], np.cumsum(avg)])
    if not cum[-1] > 0:
        return grid
    new = np.interp(np.arange(ninc + 1) * (cum[-1] / ninc), cum, grid)
    new[0], new[-1] = grid[0], grid[-1]
    return new


class Trainer:
    """Owns a bare engine configured for the training target (hydrogen, Z = A = 1, as the reference's map files)."""

    def __init__(self, Z=1.0, A=1.0, mT=1.0, mV=0.0, device=0, Eg_min=0.001, Ee_min=0.005):
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError("map training needs a CUDA device")
        cfg = capi.pb_config()
        cfg.Z_T, cfg.A_T, cfg.rho, cfg.dEdx_GeV_per_m, cfg.mT_sampler = Z, A, 1.0, 0.2, mT
        cfg.min_energy, cfg.Eg_min, cfg.Ee_min, cfg.maxF_fudge, cfg.rescale_MCS, cfg.max_sweeps = 0.01, Eg_min, Ee_min, 1.0, 1.0, 10000
        cfg.mV = mV
        self.mV, self.mT, self.Eg_min, self.Ee_min = mV, mT, Eg_min, Ee_min
        self._engine = capi.pb_engine()
        rc = capi.lib.pb_create(C.byref(self._engine), device, C.byref(cfg))
        if rc != capi.PB_OK:
            raise capi.EngineError(rc, "pb_create failed")

    def __del__(self):
        try:
            capi.lib.pb_destroy(self._engine)
        except Exception:
            pass

    def sweep(self, process, grids, ninc, E, n_points, seed, power=2.0):
        nE, stride = grids.shape
        d, n, I = np.zeros((nE, stride)), np.zeros((nE, stride)), np.zeros(nE)
        ninc32 = np.ascontiguousarray(ninc, dtype=np.int32)
        g = np.ascontiguousarray(grids, dtype=np.float64)
        E = np.ascontiguousarray(E, dtype=np.float64)
        capi.check(self._engine, capi.lib.pb_train_accumulate_p(self._engine, CODE[process], capi.dptr(g), nE, len(ninc), capi.iptr(ninc32),
                                                                capi.dptr(E), int(n_points), int(seed), float(self.mT), float(power),
                                                                capi.dptr(d), capi.dptr(n), capi.dptr(I)))
        return d, n, I

    def train(self, process, energies, nitn=None, n_points=1_000_000, alpha=1.0, seed=20261017, verbose=False, schedule=None,
              power=TRAIN_POWER):
        """-> (grids (nE, sum(ninc+1)), ninc, integral estimate of the last sweep (nE,)).  ``schedule`` = [(iterations, points per
        iteration, alpha), ...] (default ``TRAIN_SCHEDULE``: a coarse undamped stage, then a stage with more points and alpha = 1/2,
        whose training data are less noisy per increment); ``nitn`` selects the single stage (nitn, n_points, alpha) instead.
        ``power``: the training weight |jac f|^power (module docstring); 2.0 is the reference's (vegas') criterion."""
        stages = list(schedule) if schedule is not None else ([(nitn, n_points, alpha)] if nitn else list(TRAIN_SCHEDULE))
        E = np.asarray(energies, dtype=np.float64)
        dim = tb.PROC_DIM[process]
        ninc = NINC[dim]
        offs = np.concatenate([[0], np.cumsum([n + 1 for n in ninc])])
        grids = np.empty((len(E), offs[-1]))
        for k, e in enumerate(E):
            for ax, (lo, hi) in enumerate(integration_range(process, e, self.mV, self.Eg_min, self.Ee_min)):
                grids[k, offs[ax]:offs[ax + 1]] = np.linspace(lo, hi, ninc[ax] + 1)
        I = np.zeros(len(E))
        it = 0
        for nitn_s, n_points, alpha in stages:
            for _ in range(nitn_s):
                d, n, I = self.sweep(process, grids, ninc, E, n_points, seed + it, power)
                for k in range(len(E)):
                    for ax in range(dim):
                        a, b = offs[ax], offs[ax + 1]
                        grids[k, a:b] = refine(grids[k, a:b], d[k, a:b - 1], n[k, a:b - 1], alpha)
                if verbose:
                    print(process, "iteration", it, "integral[mid]", I[len(E) // 2], flush=True)
                it += 1
        _, _, I = self.sweep(process, grids, ninc, E, n_points, seed + it)
        return grids, np.array(ninc, dtype=np.int32), I


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--xsec-from", required=True, help="reference-format directory whose sm_xsec.pkl gives the energy lists")
    ap.add_argument("--out", required=True)
    ap.add_argument("--processes", default="Brem,PairProd,MuonBrem")
    ap.add_argument("--nitn", type=int, default=0, help="0: the staged default schedule (TRAIN_SCHEDULE)")
    ap.add_argument("--points", type=int, default=1_000_000)
    ap.add_argument("--power", type=float, default=TRAIN_POWER)
    a = ap.parse_args()
    import pickle
    xs = pickle.load(open(os.path.join(a.xsec_from, "sm_xsec.pkl"), "rb"))
    os.makedirs(a.out, exist_ok=True)
    tr = Trainer()
    out = {}
    for P in a.processes.split(","):
        E = np.array([row[0] for row in next(iter(xs[P].values()))])
        grids, ninc, I = tr.train(P, E, nitn=a.nitn or None, n_points=a.points, verbose=True, power=a.power)
        out[f"{P}/E"], out[f"{P}/ninc"], out[f"{P}/grid"] = E, ninc, grids
        out[f"{P}/meta"] = np.array([300, 0.001, 0.005])
        out[f"{P}/sigma_hydrogen"] = I
    np.savez_compressed(os.path.join(a.out, "sm_maps_trained.npz"), **out)


if __name__ == "__main__":
    main()
