"""max_F / sigma table builder (oracle).  TEST INFRASTRUCTURE - see oracle/__init__.py.

Restates reference ``utilities/find_maxes.py:55-119`` (do_find_max_work): for every trained map and
every target material, ``n_trials`` sweeps of ``B`` points, ``max_F = max(wgt*f)`` and
``sigma = sum(wgt*f)/n_trials``, with ``event_info`` taken from the map's training parameters and
``Z_T, A_T, mT`` from ``target_information`` (true nuclear mass here; the sampler later passes ``A_T`` as
``mT`` - quirk Q-19).

The reference's own ``sm_maps.pkl`` / ``dark_maps.pkl`` are missing upstream (.MISSING_LARGE_BLOBS), and the
procedure is stochastic, so the files written here are NEW fixtures (seeded), not copies:

    python -m oracle.findmax            # writes data/sm_maxF.npz, data/dark_maxF.npz
"""
import os
import sys

import numpy as np

from .consts import TARGETS, SM_PROCESSES, DARK_PROCESSES, m_electron, m_muon
from .integrands import DSIGMA
from .vegasmap import map_points

MATERIALS = ["graphite", "lead", "iron", "aluminum", "molybdenum"]
N_TRIALS = 100  # find_maxes.py (params["n_trials"] default)
SEED = 20261017


def split_grid(flat, ninc):
    out, off = [], 0
    for n in ninc:
        out.append(flat[off:off + n + 1])
        off += n + 1
    return out


def event_info_for(process, E, material, mV, Eg_min, Ee_min, sampler_quirk=False):
    t = TARGETS[material]
    ev = dict(E_inc=float(E), Z_T=t["Z_T"], A_T=t["A_T"], mT=(t["A_T"] if sampler_quirk else t["mT"]),
              mV=float(mV), Eg_min=Eg_min, Ee_min=Ee_min, m_lepton=m_electron)
    if process in ("MuonBrem", "DarkMuonBrem"):
        ev["m_lepton"] = m_muon
    return ev


def find_max_one(process, grid, E, material, mV, Eg_min, Ee_min, B, rng, n_trials=N_TRIALS, y=None):
    """-> (max_F, sigma_mc).  ``y``: optional pre-drawn (n_trials*B, dim) uniforms (shared across materials)."""
    if y is None:
        y = rng.random((n_trials * B, len(grid)))
    x, jac = map_points(grid, y)
    ev = event_info_for(process, E, material, mV, Eg_min, Ee_min)
    MM = (jac / B) * DSIGMA[process](x, ev)
    # np.max over a batch containing NaN is NaN and "NaN > max_F" is False in the reference: such a batch
    # never raises max_F.  Batches are B points each.
    MMb = MM.reshape(n_trials, B)
    bmax = np.max(MMb, axis=1)
    good = ~np.isnan(bmax)
    mf = float(np.max(bmax[good])) if good.any() else 0.0
    mf = max(mf, 0.0)
    return mf, float(np.sum(MM) / n_trials)


def build(maps_npz, processes, mV, rng, materials=MATERIALS, verbose=True):
    z = np.load(maps_npz)
    maxF, xs = {}, {}
    for P in processes:
        E = z[f"{P}/E"]
        ninc = z[f"{P}/ninc"]
        neval, Eg_min, Ee_min = z[f"{P}/meta"]
        B = int(neval)
        G = z[f"{P}/grid"]
        for m in materials:
            maxF[f"{P}/{m}"] = np.zeros(len(E))
            xs[f"{P}/{m}"] = np.zeros(len(E))
        for ie in range(len(E)):
            grid = split_grid(G[ie], ninc)
            y = rng.random((N_TRIALS * B, len(grid)))
            for m in materials:
                mf, sg = find_max_one(P, grid, E[ie], m, mV, Eg_min, Ee_min, B, rng, y=y)
                maxF[f"{P}/{m}"][ie] = mf
                xs[f"{P}/{m}"][ie] = sg
        if verbose:
            print(P, mV, "done", file=sys.stderr)
    return maxF, xs


def main():
    root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "data")
    rng = np.random.default_rng(SEED)
    maxF, xs = build(os.path.join(root, "sm_maps.npz"), SM_PROCESSES, 0.0, rng)
    np.savez_compressed(os.path.join(root, "sm_maxF.npz"), **maxF)
    np.savez_compressed(os.path.join(root, "sm_xsec_mc.npz"), **xs)
    dmax, dxs = {}, {}
    for f in sorted(os.listdir(root)):
        if f.startswith("dark_maps_mV") and f.endswith(".npz"):
            tag = f[len("dark_maps_mV"):-4]
            mf, sg = build(os.path.join(root, f), DARK_PROCESSES, float(tag), rng)
            dmax.update({f"{tag}/{k}": v for k, v in mf.items()})
            dxs.update({f"{tag}/{k}": v for k, v in sg.items()})
    np.savez_compressed(os.path.join(root, "dark_maxF.npz"), **dmax)
    np.savez_compressed(os.path.join(root, "dark_xsec_mc.npz"), **dxs)


if __name__ == "__main__":
    main()
