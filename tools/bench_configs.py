#!/usr/bin/env python
"""Throughput of the five BASELINE.json configurations on one GPU (SM pass + dark pass where the config has one).

    python tools/bench_configs.py [--primaries N] [--configs 1,2,3,5]

Not the contract benchmark (that is bench.py, config 2); this fills the table in DESIGN.md.  Config 4 needs the
400 GeV 4-D maps that are missing upstream and is skipped.
"""
import argparse, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from petite_b200.shower import Shower
from petite_b200.dark_shower import DarkShower
from petite_b200.constants import m_electron, m_muon

DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "data", "")
CONFIGS = {
    1: dict(name="10 GeV e- -> graphite (README example)", material="graphite", pid=11, E=10.0, mass=m_electron, mV=None, n=1000),
    2: dict(name="10 GeV gamma -> lead", material="lead", pid=22, E=10.0, mass=0.0, mV=None, n=100000),
    3: dict(name="dark: 10 GeV e- -> graphite, mV = 3 MeV (intended) ", material="graphite", pid=11, E=10.0, mass=m_electron, mV=0.003,
            active=["DarkBrem", "DarkAnn", "DarkComp"], n=100000),
    31: dict(name="dark: 10 GeV e- -> graphite, mV_in = 1 MeV -> runs as 1.0 GeV (Q-2, literal)", material="graphite", pid=11, E=10.0,
             mass=m_electron, mV=0.001, active=["DarkBrem", "DarkAnn", "DarkComp"], n=100000),
    5: dict(name="dark: 100 GeV mu- -> lead, mV = 30 MeV", material="lead", pid=13, E=100.0, mass=m_muon, mV=0.03,
            active=["DarkMuonBrem", "DarkBrem", "DarkAnn", "DarkComp"], n=100000),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--primaries", type=int, default=0)
    ap.add_argument("--configs", default="1,2,3,31,5")
    ap.add_argument("--batch", type=int, default=20000)
    a = ap.parse_args()
    for c in [int(x) for x in a.configs.split(",")]:
        cfg = CONFIGS[c]
        n = a.primaries or cfg["n"]
        if cfg["mV"] is None:
            sh = Shower(DATA, cfg["material"], 0.010, seed=20261017)
            dk = None
        else:
            sh = dk = DarkShower(DATA, cfg["material"], 0.010, cfg["mV"], active_processes=cfg["active"], seed=20261017)
        pz = np.sqrt(cfg["E"] ** 2 - cfg["mass"] ** 2)
        p = np.tile([cfg["E"], 0, 0, pz], (n, 1)); r = np.zeros((n, 3)); w = np.ones(n); m = np.full(n, cfg["mass"])
        pid = np.full(n, cfg["pid"], np.int32); fl = np.zeros(n, np.int32)
        batch = min(a.batch, n)
        sh.run_tallies(p[:batch], r[:batch], w[:batch], m[:batch], pid[:batch], fl[:batch], batch=batch, dark=dk)   # warm-up
        torch.cuda.synchronize(); t0 = time.perf_counter()
        sm_t, dk_t, tot = sh.run_tallies(p, r, w, m, pid, fl, batch=batch, dark=dk)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        out = dict(config=c, name=cfg["name"], primaries=n, seconds=dt, showers_per_s=n / dt, particle_steps_per_s=tot["n_steps"] / dt,
                   records_per_shower=tot["n_particles"] / n, waves=tot["n_waves"], trials_per_sample=tot["n_trials"] / max(tot["n_samples"], 1))
        if dk is not None:
            t = dk_t.cpu().numpy()
            out.update(dark_vectors_per_shower=tot["n_dark"] / n, dark_weight_sum_per_shower=float(t[8 + 5] / n), mV_used=dk._mV)
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
