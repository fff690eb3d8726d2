#!/usr/bin/env python
"""Per-kernel CUDA-event times of the DARK pass (pb_run_dark) for BASELINE configs 3 and 5: 2e4 / 1e4 SM showers, then
generate_dark_showers with profiling on.  PB_SAMPLE_T_DARK / PB_SAMPLE_G pass through.   python tools/dark_profile.py [3 5]"""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from petite_b200.dark_shower import DarkShower
from petite_b200.constants import m_electron, m_muon
DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "data", "")
CFG = {3: ("graphite", 11, 10.0, m_electron, 0.003, ["DarkBrem", "DarkAnn", "DarkComp"], 20000),
       5: ("lead", 13, 100.0, m_muon, 0.030, ["DarkMuonBrem", "DarkBrem", "DarkAnn", "DarkComp"], 5000)}
for c in [int(a) for a in sys.argv[1:]] or [3, 5]:
    mat, pid, E, m, mV, act, n = CFG[c]
    sh = DarkShower(DATA, mat, 0.010, mV, active_processes=act, seed=20261017)
    p = np.tile([E, 0, 0, np.sqrt(E * E - m * m)], (n, 1))
    b = sh.run_arrays(p, np.zeros((n, 3)), np.ones(n), np.full(n, m), np.full(n, pid, np.int32), np.zeros(n, np.int32), first_shower_id=0)
    d = sh.generate_dark_showers(b)            # warm-up
    sh.set_profiling(2)
    best = None
    for rep in range(3):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); d = sh.generate_dark_showers(b); e1.record(); torch.cuda.synchronize()
        pr = sh.get_profile()
        row = {"config": c, "sm_showers": n, "sm_records": b.n, "dark_vectors": d.n, "dark_trials": d.counters["n_trials"], "dark_samples": d.counters["n_samples"],
               "pass_ms": e0.elapsed_time(e1), "kernels_ms": {k: round(v, 3) for k, v in pr["ms"].items() if v},
               "trials": {k: v for k, v in pr["trials"].items() if v}, "samples": {k: v for k, v in pr["samples"].items() if v},
               "T_dark": os.environ.get("PB_SAMPLE_T_DARK", "default"), "G": os.environ.get("PB_SAMPLE_G_DARK", "default")}
        if best is None or row["pass_ms"] < best["pass_ms"]:
            best = row
    best["ps_per_trial"] = 1e9 * best["kernels_ms"].get("k_sample", 0.0) / max(best["dark_trials"], 1)
    print(json.dumps(best), flush=True)
    del sh, b, d
    torch.cuda.empty_cache()
