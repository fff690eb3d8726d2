import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from tests.gpu_util import probe
from tests.test_gpu_dark import dark_shower
from petite_b200 import _capi as capi
g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "integrands.npz"))
for material in ("graphite", "lead"):
    ds = dark_shower(material, 0.03)
    for process, code in (("DarkBrem", 8), ("DarkMuonBrem", 11)):
        E, x, f = g[f"{material}/{process}/E"], g[f"{material}/{process}/x"], g[f"{material}/{process}/f"]
        for off in (0, 32):
            got = probe(ds, capi.PROBE_DSIGMA, code + off, np.column_stack([E, x]), 1)[:, 0]
            nz = f != 0
            rel = np.abs(got[nz] - f[nz]) / np.abs(f[nz])
            zero_mismatch = int(np.sum((got == 0) != (f == 0)))
            worst = np.argsort(-rel)[:3]
            print(material, process, "fast" if off else "exact", "zero mismatch", zero_mismatch, "median rel", np.median(rel), "max rel", rel.max(),
                  [(float(E[nz][i]), x[nz][i][:3].tolist(), float(f[nz][i]), float(got[nz][i])) for i in worst[:2]])
print("---- violations of the test bound")
for material in ("graphite", "lead"):
    ds = dark_shower(material, 0.03)
    for process, code in (("DarkBrem", 8), ("DarkMuonBrem", 11)):
        E, x, f = g[f"{material}/{process}/E"], g[f"{material}/{process}/x"], g[f"{material}/{process}/f"]
        cond = 4.4e-16 / 10.0 ** x[:, 1]
        got = probe(ds, capi.PROBE_DSIGMA, code, np.column_stack([E, x]), 1)[:, 0]
        for Einc in np.unique(E):
            sel = E == Einc
            scale = np.max(np.abs(f[sel])) if np.any(f[sel] != 0) else 1.0
            tol = (1e-12 + 1e-11 + cond[sel]) * np.abs(f[sel]) + 1e-9 * scale
            bad = np.abs(got[sel] - f[sel]) > tol
            for i in np.where(bad)[0]:
                print(material, process, Einc, x[sel][i][:3].tolist(), f[sel][i], got[sel][i], abs(got[sel][i] - f[sel][i]) / abs(f[sel][i]), cond[sel][i], scale)
import time, torch
from tests.gpu_util import primaries
ds = dark_shower("graphite", 0.003) if False else None
from petite_b200.dark_shower import DarkShower
from tests.conftest import DATA
d3 = DarkShower(DATA, "graphite", 0.010, 0.003, seed=1)
sm = d3.generate_showers(primaries(11, 10.0, 20000), first_shower_id=0)
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    dk = d3.generate_dark_showers(sm)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    print("dark pass of 2e4 showers: %.1f ms" % ((t1 - t0) * 1e3), dk.n, dk.counters["n_trials"])
