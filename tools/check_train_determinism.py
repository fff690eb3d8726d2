"""Is map training reproducible?  Train the same rows twice (and once with a single CUDA block per row, where the fp64 atomics of a
block still run in any order) and compare the grids; find_max twice on the same maps."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from petite_b200.train import Trainer
from petite_b200.shower import Shower
DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "data", "")
xs = np.load(DATA + "sm_xsec.npz")["Brem/graphite"]
E = xs[[60, 80, 99], 0]
for power in (2.0, 8.0):
    tr = Trainer()
    g = [tr.train("Brem", E, power=power)[0] for _ in range(3)]
    print("power", power, "max |grid difference| between identical trainings:", [float(np.max(np.abs(g[0] - x))) for x in g[1:]],
          "after ONE iteration:", float(np.max(np.abs(tr.train("Brem", E, power=power, nitn=1)[0] - tr.train("Brem", E, power=power, nitn=1)[0]))))
    d1 = tr.sweep("Brem", g[0], [960, 1000, 1000, 1000], E, 2_000_000, 5, power)
    d2 = tr.sweep("Brem", g[0], [960, 1000, 1000, 1000], E, 2_000_000, 5, power)
    rel = np.abs(d1[0] - d2[0]) / np.maximum(np.abs(d1[0]), 1e-300)
    print("   one sweep twice: max relative difference of the training sums", float(rel.max()), "counts equal", bool(np.array_equal(d1[1], d2[1])))
sh = Shower(DATA, "graphite", 0.010, seed=3)
a = sh.find_max("Brem", n_trials=400, seed=9); b = sh.find_max("Brem", n_trials=400, seed=9)
print("find_max twice: identical", bool(np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])))
