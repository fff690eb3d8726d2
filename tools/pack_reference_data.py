#!/usr/bin/env python
"""Repack PETITE's on-disk tables into this project's pickle-free ``.npz`` layout.

Reads a reference-format directory (``<Proc>/<Proc>_AdaptiveMaps.npy``, ``sm_xsec.pkl``,
``dark_xsec.pkl``, ``dark_weights.pkl``, ``dark_drate.pkl``; formats per SURVEY.md 3.5) and
writes

    sm_maps.npz            <P>/E (nE,)  <P>/ninc (dim,)  <P>/grid (nE, sum(ninc+1))
                           <P>/meta = [neval, Eg_min, Ee_min]
    sm_xsec.npz            <P>/<material> (n,2)
    dark_maps_mV<m>.npz    same layout as sm_maps, one file per trained mass
    dark_xsec.npz          <mV>/<P>/<material> (n,2)
    dark_weights.npz       <mV>/<material>/<name> (n,2)
    dark_drate.npz         <mV>/<material>/<name>/E (n,)  .../table (n,10,2)

The ``.npy`` map files are pickles of ``vegas._vegas.AdaptiveMap``; they reduce to a
list-of-lists of node positions and are read with a stub class (no vegas needed).
max_F tables are NOT produced here (the reference's ``sm_maps.pkl``/``dark_maps.pkl`` are
missing upstream); see ``python -m oracle.findmax``.
"""
import argparse
import os
import pickle
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from petite_b200.tables import load_adaptive_maps_npy, SM_PROCESSES, DARK_PROCESSES, mv_tag  # noqa: E402

NEVAL_DEFAULT = 300  # utilities/generate_integrators.py:206-210, SURVEY.md 3.5


def pack_maps(rows):
    E = np.array([float(p["E_inc"]) for p, _ in rows])
    ninc = np.array([len(g) - 1 for g in rows[0][1]], dtype=np.int32)
    grid = np.stack([np.concatenate([np.asarray(g, dtype=np.float64) for g in grid]) for _, grid in rows])
    p0 = rows[0][0]
    meta = np.array([NEVAL_DEFAULT, p0.get("Eg_min", 0.001), p0.get("Ee_min", 0.005)], dtype=np.float64)
    return {"E": E, "ninc": ninc, "grid": grid, "meta": meta}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--src", default="/root/reference/data/")
    ap.add_argument("--dst", default=os.path.join(os.path.dirname(__file__), "..", "data"))
    ap.add_argument("--masses", default="0.003,0.03,1.0", help="dark masses (GeV) to pack maps for")
    ap.add_argument("--skip-missing", action="store_true",
                    help="pack only the map files that exist (data_400GeV ships the 1-D maps only; the 4-D/3-D ones are "
                         "retrained by tools/make_400GeV.py)")
    a = ap.parse_args()
    os.makedirs(a.dst, exist_ok=True)

    out = {}
    for P in SM_PROCESSES:
        path = os.path.join(a.src, P, f"{P}_AdaptiveMaps.npy")
        if a.skip_missing and not os.path.exists(path):
            continue
        rows = load_adaptive_maps_npy(path)
        for k, v in pack_maps(rows).items():
            out[f"{P}/{k}"] = v
    np.savez_compressed(os.path.join(a.dst, "sm_maps_shipped.npz" if a.skip_missing else "sm_maps.npz"), **out)

    xs = pickle.load(open(os.path.join(a.src, "sm_xsec.pkl"), "rb"))
    np.savez_compressed(os.path.join(a.dst, "sm_xsec.npz"),
                        **{f"{P}/{m}": np.asarray(v, dtype=np.float64) for P, d in xs.items() for m, v in d.items()})

    dx = pickle.load(open(os.path.join(a.src, "dark_xsec.pkl"), "rb"))
    np.savez_compressed(os.path.join(a.dst, "dark_xsec.npz"),
                        **{f"{mv_tag(mV)}/{P}/{m}": np.asarray(v, dtype=np.float64)
                           for mV, d in dx.items() for P, dd in d.items() for m, v in dd.items()})

    for mV in [float(s) for s in a.masses.split(",") if s]:
        out = {}
        for P in DARK_PROCESSES:
            path = os.path.join(a.src, P, f"mV_{int(round(mV * 1000))}MeV", f"{P}_AdaptiveMaps.npy")
            if a.skip_missing and not os.path.exists(path):
                continue
            rows = load_adaptive_maps_npy(path)
            for k, v in pack_maps(rows).items():
                out[f"{P}/{k}"] = v
        np.savez_compressed(os.path.join(a.dst, f"dark_maps_mV{mv_tag(mV)}" + ("_shipped" if a.skip_missing else "") + ".npz"), **out)

    w = pickle.load(open(os.path.join(a.src, "dark_weights.pkl"), "rb"))
    np.savez_compressed(os.path.join(a.dst, "dark_weights.npz"),
                        **{f"{mv_tag(mV)}/{m}/{n}": np.asarray(v, dtype=np.float64)
                           for mV, d in w.items() for m, dd in d.items() for n, v in dd.items()})
    r = pickle.load(open(os.path.join(a.src, "dark_drate.pkl"), "rb"))
    out = {}
    for mV, d in r.items():
        for m, dd in d.items():
            for n, tab in dd.items():
                keys = list(tab.keys())
                out[f"{mv_tag(mV)}/{m}/{n}/E"] = np.array([float(k) for k in keys])
                out[f"{mv_tag(mV)}/{m}/{n}/table"] = np.stack([np.asarray(tab[k], dtype=np.float64) for k in keys])
    np.savez_compressed(os.path.join(a.dst, "dark_drate.npz"), **out)
    for f in sorted(os.listdir(a.dst)):
        print(f, os.path.getsize(os.path.join(a.dst, f)))


if __name__ == "__main__":
    main()
