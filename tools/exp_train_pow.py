#!/usr/bin/env python
"""Training-weight experiment: VEGAS refines its map on sum (jac f)^2 per increment; with vegas' adaptive stratified
sampling (the reference trains with nstrat 60 x 50 x 50 x 50, beta 0.75) a hypercube holding n_h ~ sigma_h^beta samples contributes
(jac f)^2 / n_h.  Trainer.train(power=p) trains on |jac f|^p (pb_train_accumulate_p); this tool
reports the accept/reject efficiency sigma / (B max_F) of maps trained with several p relative to the shipped maps.
    python tools/exp_train_pow.py Brem PairProd"""
import sys, time, json, os
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from petite_b200.train import Trainer
from petite_b200 import tables as tb
from petite_b200.shower import Shower, process_code
DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "data", "")
SCHED = [(10, 2_000_000, 1.0), (10, 8_000_000, 0.5)]
NT = int(os.environ.get("EXP_TRIALS", "2000"))
for P in sys.argv[1:] or ["Brem"]:
    xs = np.load(DATA + "sm_xsec.npz")[f"{P}/graphite"]
    rows = [20, 40, 60, 70, 80, 90, 99]; E = xs[rows, 0]
    sh = Shower(DATA, "graphite", 0.010, seed=3)
    shipped = sh._maps[P]
    mf_old, sg_old = sh.find_max(P, n_trials=NT, seed=9)
    eff_old = (sg_old / (300 * mf_old))[rows]
    print(json.dumps({"process": P, "E": E.round(3).tolist(), "shipped_eff": eff_old.round(4).tolist()}), flush=True)
    for power in [float(v) for v in os.environ.get("EXP_POWERS", "2,1.5,1.25,1,3").split(",")]:
        tr = Trainer(); t = time.time()
        grids, ninc, I = tr.train(P, E, schedule=SCHED, power=power)
        dt = time.time() - t
        ms = tb.MapSet(P, E, ninc, grids, np.ones(len(E)), shipped.neval, shipped.Eg_min, shipped.Ee_min)
        sh._upload_maps(process_code[P], ms); sh._maps[P] = ms
        mf, sg = sh.find_max(P, n_trials=NT, seed=9)
        r = sg / (300 * mf) / eff_old
        print(json.dumps({"process": P, "power": power, "seconds": round(dt, 1), "sigma_ratio": (sg / xs[rows, 1]).round(4).tolist(),
                          "eff_over_shipped": r.round(3).tolist(), "geo_mean": round(float(np.exp(np.mean(np.log(r)))), 3)}), flush=True)
        del tr
