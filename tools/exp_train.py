#!/usr/bin/env python
"""Training-schedule experiment: accept/reject efficiency sigma / (B max_F) of maps trained here relative to the
shipped maps, for a few energies of one process.   python tools/exp_train.py Brem PairProd"""
import sys, time, json, os
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from petite_b200.train import Trainer
from petite_b200 import tables as tb
from petite_b200.shower import Shower, process_code
DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "data", "")
SCHEDULES = {
    "r1 (30 x 1e6, alpha 1)": [(30, 1_000_000, 1.0)],
    "10x1e6 a1 + 10x4e6 a.5": [(10, 1_000_000, 1.0), (10, 4_000_000, 0.5)],
    "10x1e6 a1 + 10x4e6 a.5 + 5x1.6e7 a.25": [(10, 1_000_000, 1.0), (10, 4_000_000, 0.5), (5, 16_000_000, 0.25)],
    "10x2e6 a.5 + 10x1.6e7 a.5": [(10, 2_000_000, 0.5), (10, 16_000_000, 0.5)],
    "vegas-like 20 x 1.2e7 a.5": [(20, 12_000_000, 0.5)],
}
for P in sys.argv[1:] or ["Brem"]:
    xs = np.load(DATA + "sm_xsec.npz")[f"{P}/graphite"]
    rows = [30, 60, 80, 99]; E = xs[rows, 0]
    sh = Shower(DATA, "graphite", 0.010, seed=3)
    shipped = sh._maps[P]
    mf_old, sg_old = sh.find_max(P, n_trials=400, seed=9)
    eff_old = (sg_old / (300 * mf_old))[rows]
    print(json.dumps({"process": P, "shipped_eff": eff_old.tolist()}), flush=True)
    for name, sched in SCHEDULES.items():
        tr = Trainer(); t = time.time()
        grids, ninc, I = tr.train(P, E, schedule=sched)
        dt = time.time() - t
        ms = tb.MapSet(P, E, ninc, grids, np.ones(len(E)), shipped.neval, shipped.Eg_min, shipped.Ee_min)
        sh._upload_maps(process_code[P], ms); sh._maps[P] = ms
        mf, sg = sh.find_max(P, n_trials=400, seed=9)
        print(json.dumps({"process": P, "schedule": name, "seconds_4_energies": round(dt, 1), "sigma_ratio": (sg / xs[rows, 1]).round(4).tolist(),
                          "eff_over_shipped": (sg / (300 * mf) / eff_old).round(3).tolist()}), flush=True)
        del tr
