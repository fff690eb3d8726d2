"""Distribution of accept/reject trials per sample (config 2), overall and by the parent's energy decade."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from petite_b200.shower import Shower
DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "data", "")
n = 4000
sh = Shower(DATA, "lead", 0.010, seed=20261017)
p = np.tile([10.0, 0, 0, 10.0], (n, 1))
b = sh.run_arrays(p, np.zeros((n, 3)), np.ones(n), np.zeros(n), np.full(n, 22, dtype=np.int32), np.zeros(n, dtype=np.int32))
h = b.to_host()
nt = h["ntrials"]; E = h["pf"][:, 0]; pid = h["pid"]
m = nt > 0
print("samples", m.sum(), "mean trials", nt[m].mean(), "p50/p90/p99/p99.9/max", np.percentile(nt[m], [50, 90, 99, 99.9]), nt[m].max())
tot = nt[m].sum()
for k in (50, 100, 200, 500, 1000):
    print(f"  samples with > {k} trials: {np.mean(nt[m] > k):.2e} of samples, {nt[m][nt[m] > k].sum() / tot:.3f} of all trials")
for name, sel in (("photon", pid == 22), ("e-", pid == 11), ("e+", pid == -11)):
    print(name)
    edges = np.logspace(-2, 1, 13)
    for lo, hi in zip(edges[:-1], edges[1:]):
        s = m & sel & (E >= lo) & (E < hi)
        if s.sum() > 20:
            print(f"   E in [{lo:.3g}, {hi:.3g}): n {s.sum():8d} mean {nt[s].mean():7.1f} p99 {np.percentile(nt[s], 99):7.0f} max {nt[s].max():6d}  share of trials {nt[s].sum() / tot:.3f}")
