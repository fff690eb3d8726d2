#!/usr/bin/env python
"""Oracle-side ensembles for the >= 1e5-shower statistical tests (tests/test_gpu_ensemble.py).

    python tests/golden/make_ensemble.py [config ...]   # -> tests/golden/ensemble.npz  (about 20 minutes on 8 cores)

For each configuration the CPU oracle (counter-mode draws, shower ids well away from the ones the GPU test uses) steps
N independent showers and keeps PER-SHOWER summaries only - showers are the independent units a KS / chi-square test may
pool, the particles inside one shower are not.  The same summaries are computed from the GPU stack by
``tests/ensemble_stats.py``; both sides share the definitions in that module.
"""
import os
import sys
from multiprocessing import Pool

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests.ensemble_stats import CONFIGS, summarise_oracle_sm, summarise_oracle_dark   # noqa: E402

FIRST_ID = 5_000_000        # the GPU test uses ids from 0


def _one(args):
    name, i = args
    cfg = CONFIGS[name]
    from oracle.shower import OParticle
    if cfg.get("mV") is None:
        from oracle.shower import OracleShower
        o = OracleShower(None, cfg["material"], cfg["E_min"], seed=cfg["seed"], rng="counter")
    else:
        from oracle.dark import OracleDarkShower
        o = OracleDarkShower(None, cfg["material"], cfg["E_min"], cfg["mV"], seed=cfg["seed"], rng="counter")
    E, m = cfg["E0"], cfg["mass"]
    p = OParticle([E, 0.0, 0.0, float(np.sqrt(E * E - m * m))], (0.0, 0.0, 0.0), PID=cfg["pid"], ID=1, mass=m)
    sm = o.generate_shower(p, shower_id=FIRST_ID + i)
    row = summarise_oracle_sm(sm)
    if cfg.get("mV") is not None:
        _, vs = o.generate_dark_shower(sm)
        row.update(summarise_oracle_dark(vs))
    return row


def main():
    path = os.path.join(ROOT, "tests", "golden", "ensemble.npz")
    only = [a for a in sys.argv[1:] if a in CONFIGS]            # e.g. `make_ensemble.py c5_mu_lead`: redo these, keep the rest
    out = dict(np.load(path)) if only and os.path.exists(path) else {}
    with Pool(os.cpu_count()) as pool:
        for name, cfg in CONFIGS.items():
            if only and name not in only:
                continue
            rows = pool.map(_one, [(name, i) for i in range(cfg["n_oracle"])], chunksize=4)
            for k in rows[0]:
                out[f"{name}/{k}"] = np.array([r[k] for r in rows])
            print(name, len(rows), {k: float(np.mean(out[f"{name}/{k}"])) for k in rows[0] if np.ndim(rows[0][k]) == 0})
    np.savez_compressed(path, **out)


if __name__ == "__main__":
    main()
