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
DATA400 = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "data_400GeV", "")


def beam_dump_spectrum(n, seed=20261017):
    """Config 4: 50 % photons resampled from the reference's 120 GeV pi0-photon beam scaled x(400/120) in
    momentum, 25 % e-, 25 % e+ with dN/dE ~ 1/E on [1, 400] GeV along +z."""
    rng = np.random.default_rng(seed)
    beam = np.load(DATA400 + "Photons_From_Pi0s_120GeV.npy")
    ng = n // 2
    g = beam[rng.integers(0, len(beam), ng)] * (400.0 / 120.0)
    g = g[g[:, 0] > 0.0016]
    ng = len(g)
    ne = n - ng
    E = np.exp(rng.uniform(np.log(1.0), np.log(400.0), ne))
    pe = np.column_stack([E, np.zeros(ne), np.zeros(ne), np.sqrt(E ** 2 - m_electron ** 2)])
    p = np.vstack([g, pe])
    pid = np.concatenate([np.full(ng, 22), np.where(np.arange(ne) % 2 == 0, 11, -11)]).astype(np.int32)
    m = np.where(pid == 22, 0.0, m_electron)
    return p, pid, m
CONFIGS = {
    1: dict(name="10 GeV e- -> graphite (README example)", material="graphite", pid=11, E=10.0, mass=m_electron, mV=None, n=1000),
    2: dict(name="10 GeV gamma -> lead", material="lead", pid=22, E=10.0, mass=0.0, mV=None, n=100000),
    3: dict(name="dark: 10 GeV e- -> graphite, mV = 3 MeV (intended) ", material="graphite", pid=11, E=10.0, mass=m_electron, mV=0.003,
            active=["DarkBrem", "DarkAnn", "DarkComp"], n=100000),
    31: dict(name="dark: 10 GeV e- -> graphite, mV_in = 1 MeV -> runs as 1.0 GeV (Q-2, literal)", material="graphite", pid=11, E=10.0,
             mass=m_electron, mV=0.001, active=["DarkBrem", "DarkAnn", "DarkComp"], n=100000),
    4: dict(name="dark: 400 GeV beam-dump e+-/gamma spectrum -> lead, mV = 10 MeV (retrained 400 GeV maps)", material="lead", pid=0, E=0.0,
            mass=0.0, mV=0.01, active=["DarkBrem", "DarkAnn", "DarkComp"], n=100000, data=DATA400),
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
        data = cfg.get("data", DATA)
        if cfg["mV"] is None:
            sh = Shower(data, cfg["material"], 0.010, seed=20261017)
            dk = None
        else:
            sh = dk = DarkShower(data, cfg["material"], 0.010, cfg["mV"], active_processes=cfg["active"], seed=20261017)
        if c == 4:
            p, pid, m = beam_dump_spectrum(n)
            n = len(pid)
            r = np.zeros((n, 3)); w = np.ones(n); fl = np.zeros(n, np.int32)
        else:
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
