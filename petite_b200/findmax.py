"""max_F / sigma table builder on the GPU (``pb_find_max``): the reference's ``utilities/find_maxes.py`` step.

    python -m petite_b200.findmax --dict-dir data/ --materials graphite,lead --out data/sm_maxF_gpu.npz

The reference's ``sm_maps.pkl`` / ``dark_maps.pkl`` (which carry max_F) are missing upstream, so these tables have to be
regenerated; the procedure is stochastic (max over random sweeps), the result is a fixture, not a copy.
"""
import argparse

import numpy as np

from . import tables as tb


def build_sm(dict_dir, materials, n_trials=100, seed=20261017):
    from .shower import Shower
    out_max, out_sig = {}, {}
    for m in materials:
        sh = Shower(dict_dir, m, 0.010)
        for P in tb.SM_PROCESSES:
            mf, sg = sh.find_max(P, n_trials=n_trials, seed=seed)
            out_max[f"{P}/{m}"], out_sig[f"{P}/{m}"] = mf, sg
    return out_max, out_sig


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dict-dir", default="data/")
    ap.add_argument("--materials", default="graphite,lead,iron,aluminum,molybdenum")
    ap.add_argument("--out", default="sm_maxF_gpu.npz")
    ap.add_argument("--n-trials", type=int, default=100)
    a = ap.parse_args()
    mf, sg = build_sm(a.dict_dir, a.materials.split(","), a.n_trials)
    np.savez_compressed(a.out, **mf)
    np.savez_compressed(a.out.replace(".npz", "_sigma.npz"), **sg)


if __name__ == "__main__":
    main()
