#!/usr/bin/env python
"""Build the 400 GeV table set that BASELINE config 4 needs (run on a GPU box).

Upstream ships data_400GeV/ with cross-sections and the 1-D maps only; the 4-D (Brem, PairProd, MuonBrem) and 3-D
(DarkBrem, DarkMuonBrem) map files are missing (.MISSING_LARGE_BLOBS).  This retrains them with petite_b200.train on
the reference's own 150-energy grids, rebuilds max_F for the chosen materials with the GPU find_max, and writes
``data_400GeV/{sm_maps,sm_maxF,dark_maps_mV<m>,dark_maxF}.npz`` next to the repacked shipped files.

    python tools/make_400GeV.py [--materials lead] [--mV 0.01]
"""
import argparse, os, sys, time
import numpy as np
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from petite_b200 import tables as tb
from petite_b200.train import Trainer, TRAIN_POWER


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dir", default=os.path.join(ROOT, "data_400GeV", ""))
    ap.add_argument("--materials", default="lead")
    ap.add_argument("--mV", type=float, default=0.01)
    ap.add_argument("--nitn", type=int, default=0, help="0: petite_b200.train.TRAIN_SCHEDULE")
    ap.add_argument("--points", type=int, default=1_000_000)
    ap.add_argument("--power", type=float, default=None, help="training weight |jac f|^p (default petite_b200.train.TRAIN_POWER = 8)")
    a = ap.parse_args()
    power = TRAIN_POWER if a.power is None else a.power
    D = a.dir
    mats = a.materials.split(",")
    t0 = time.time()
    xs = np.load(D + "sm_xsec.npz")
    sm = dict(np.load(D + "sm_maps_shipped.npz"))
    tr = Trainer()
    sig_h = {}
    for P in ("Brem", "PairProd", "MuonBrem"):
        E = xs[f"{P}/lead"][:, 0]
        grids, ninc, I = tr.train(P, E, nitn=a.nitn or None, n_points=a.points, power=power)
        sm[f"{P}/E"], sm[f"{P}/ninc"], sm[f"{P}/grid"], sm[f"{P}/meta"] = E, ninc, grids, np.array([300, 0.001, 0.005])
        sig_h[P] = I
        print(f"trained {P}: {len(E)} energies, {time.time() - t0:.1f} s", flush=True)
    np.savez_compressed(D + "sm_maps.npz", **sm)
    os.environ["PETITE_B200_ALLOW_MISSING_MAXF"] = "1"
    from petite_b200.shower import Shower
    from petite_b200.dark_shower import DarkShower
    mf_out, ratio = {}, {}
    for m in mats:
        sh = Shower(D, m, 0.010)
        for P in tb.SM_PROCESSES:
            mf, sg = sh.find_max(P, n_trials=100)
            mf_out[f"{P}/{m}"] = mf
            ref = xs[f"{P}/{m}"][:, 1]
            ok = ref > 0
            ratio[f"{P}/{m}"] = float(np.mean(sg[ok] / ref[ok]))
    np.savez_compressed(D + "sm_maxF.npz", **mf_out)
    print("sigma(new maps) / sigma(shipped sm_xsec), mean over 150 energies:", ratio, flush=True)
    # dark sector
    tag = tb.mv_tag(a.mV)
    dxs = np.load(D + "dark_xsec.npz")
    dk = dict(np.load(D + f"dark_maps_mV{tag}_shipped.npz"))
    trd = Trainer(mT=200.0, mV=a.mV)                       # the reference's training target: hydrogen with mT = 200 GeV
    for P in ("DarkBrem", "DarkMuonBrem"):
        E = dxs[f"{tag}/{P}/lead"][:, 0]
        grids, ninc, I = trd.train(P, E, nitn=a.nitn or None, n_points=a.points, power=power)
        dk[f"{P}/E"], dk[f"{P}/ninc"], dk[f"{P}/grid"], dk[f"{P}/meta"] = E, ninc, grids, np.array([300, 0.001, 0.005])
        print(f"trained {P}: {len(E)} energies, {time.time() - t0:.1f} s", flush=True)
    np.savez_compressed(D + f"dark_maps_mV{tag}.npz", **dk)
    dmf, dratio = {}, {}
    for m in mats:
        ds = DarkShower(D, m, 0.010, a.mV, active_processes=["DarkBrem", "DarkAnn", "DarkComp", "DarkMuonBrem"])
        for P in tb.DARK_PROCESSES:
            mf, sg = ds.find_max(P, n_trials=100)
            dmf[f"{tag}/{P}/{m}"] = mf
            ref = dxs[f"{tag}/{P}/{m}"][:, 1]
            ok = (ref > 0) & (sg > 0)
            dratio[f"{P}/{m}"] = float(np.median(sg[ok] / ref[ok])) if ok.any() else None
    np.savez_compressed(D + "dark_maxF.npz", **dmf)
    print("dark sigma(new maps) / shipped dark_xsec, median:", dratio, flush=True)
    print(f"done in {time.time() - t0:.1f} s")


if __name__ == "__main__":
    main()
