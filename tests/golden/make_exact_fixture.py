#!/usr/bin/env python
"""Oracle (counter mode) per-shower decision summaries for the EXACT-equality test at 1e3 / 1e4 showers.

    python tests/golden/make_exact_fixture.py [--workers N]     # -> tests/golden/exact_showers.npz   (~45 min on 8 cores)

For BASELINE configs 1 (1e3 x 10 GeV e- -> graphite: the whole configuration) and 2 (the first 1e4 of its 1e5 x 10 GeV gamma ->
lead) the CPU oracle steps shower ids 0 .. n-1 with the engine's draw protocol and keeps, per shower, integers only:
multiplicity, sum of accept/reject trials, sum of dE/dx sub-steps, and the histogram of generation processes (16 codes).
tests/test_gpu_showers.py::test_exact_decisions_* demands EXACT equality of all of them on the GPU: any hard-scatter, process or
accept decision that flips because the hot-loop forms (fast_rcp, hot_log, folded integrands, summed n*sigma tables) differ from
the oracle's libm path in the last bits would change at least one of these integers.  The test reports the flip rate.
"""
import os
import sys
import time
from multiprocessing import Pool

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests.ensemble_stats import CONFIGS   # noqa: E402

CASES = {"c1_e_graphite": 1_000, "c2_gamma_lead": 10_000}
_O = {}


def _one(args):
    name, i = args
    cfg = CONFIGS[name]
    from oracle.shower import OracleShower, OParticle
    from oracle.consts import PROC_CODE
    if name not in _O:
        _O[name] = OracleShower(None, cfg["material"], cfg["E_min"], seed=cfg["seed"], rng="counter")
    E, m = cfg["E0"], cfg["mass"]
    sh = _O[name].generate_shower(OParticle([E, 0.0, 0.0, float(np.sqrt(E * E - m * m))], (0.0, 0.0, 0.0), PID=cfg["pid"], ID=1, mass=m), shower_id=i)
    hist = np.zeros(16, dtype=np.int64)
    for q in sh:
        hist[PROC_CODE[q.process]] += 1
    return len(sh), sum(q.ntrials for q in sh), sum(q.nsub for q in sh), hist


def main():
    workers = int(sys.argv[sys.argv.index("--workers") + 1]) if "--workers" in sys.argv else os.cpu_count()
    path = os.path.join(ROOT, "tests", "golden", "exact_showers.npz")
    out = dict(np.load(path)) if os.path.exists(path) and "--fresh" not in sys.argv else {}     # finished cases are kept
    with Pool(workers) as pool:
        for name, n in CASES.items():
            if f"{name}/mult" in out and len(out[f"{name}/mult"]) == n:
                continue
            t0 = time.time()
            rows = pool.map(_one, [(name, i) for i in range(n)], chunksize=8)
            out[f"{name}/mult"] = np.array([r[0] for r in rows], dtype=np.int64)
            out[f"{name}/ntrials"] = np.array([r[1] for r in rows], dtype=np.int64)
            out[f"{name}/nsub"] = np.array([r[2] for r in rows], dtype=np.int64)
            out[f"{name}/proc_hist"] = np.stack([r[3] for r in rows])
            print(name, n, f"{time.time() - t0:.0f} s", "mean multiplicity", out[f"{name}/mult"].mean(), flush=True)
            np.savez_compressed(path, **out)


if __name__ == "__main__":
    main()
