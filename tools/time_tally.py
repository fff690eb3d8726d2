"""Time pb_tally on a resident batch (CUDA events)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from petite_b200.shower import Shower
DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "data", "")
n = 50_000
sh = Shower(DATA, "lead", 0.010, seed=1)
p = np.tile([10.0, 0, 0, 10.0], (n, 1))
b = sh.run_arrays(p, np.zeros((n, 3)), np.ones(n), np.zeros(n), np.full(n, 22, dtype=np.int32), np.zeros(n, dtype=np.int32), capacity=n * 1800)
t = sh.tally(b)
ref = t.cpu().numpy().copy()
a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for rep in range(3):
    t.zero_(); torch.cuda.synchronize(); a.record(); sh.tally(b, t); e.record(); torch.cuda.synchronize()
    print("records", b.n, "tally ms", a.elapsed_time(e), "GB/s (80 B/record)", b.n * 80 / a.elapsed_time(e) / 1e6)
print("same as first", np.allclose(ref, t.cpu().numpy(), rtol=1e-12))
