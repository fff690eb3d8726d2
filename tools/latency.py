#!/usr/bin/env python
"""Small-batch regime: BASELINE config 1 (1e3 x 10 GeV e- -> graphite) throughput and the single-primary latency of the drop-in
call ``Shower.generate_shower(p0)`` (the reference's own call pattern: 2.59 s per 10 GeV shower, examples/tutorial.ipynb:[37]).
Run once with PB_GRAPH=1 (default: the wave loop is one CUDA-graph launch) and once with PB_GRAPH=0 (stream launches + a host
synchronisation per growing wave).     python tools/latency.py > gpurun_out/latency.json"""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from petite_b200 import Particle
from petite_b200.shower import Shower
from petite_b200.constants import m_electron

DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "data", "")
sh = Shower(DATA, "graphite", 0.010, seed=20261017)
out = {"PB_GRAPH": os.environ.get("PB_GRAPH", "1")}
for n in (1, 10, 100, 1000, 10000):
    E = 10.0
    p = np.tile([E, 0, 0, np.sqrt(E * E - m_electron ** 2)], (n, 1))
    a = [p, np.zeros((n, 3)), np.ones(n), np.full(n, m_electron), np.full(n, 11, np.int32), np.zeros(n, np.int32)]
    cap = int(n * 1500 + 65536)
    for _ in range(3):
        b = sh.run_arrays(*a, capacity=cap, first_shower_id=0)
    reps = 20 if n <= 1000 else 5
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for k in range(reps):
        b = sh.run_arrays(*a, capacity=cap, first_shower_id=k * n)
        t = sh.tally(b)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / reps
    out[f"batch_{n}"] = {"ms": 1e3 * dt, "showers_per_s": n / dt, "waves": b.counters["n_waves"], "records": b.n}
# the drop-in call: Particle in, list of Particle out (includes building ~700 Python Particle objects)
p0 = Particle([10.0, 0, 0, np.sqrt(100 - m_electron ** 2)], [0, 0, 0], {"PID": 11, "ID": 1, "mass": m_electron})
for _ in range(3):
    lst = sh.generate_shower(p0)
ts = []
for k in range(20):
    t0 = time.perf_counter(); lst = sh.generate_shower(p0); ts.append(time.perf_counter() - t0)
out["generate_shower_p0"] = {"median_ms": 1e3 * float(np.median(ts)), "min_ms": 1e3 * min(ts), "particles": len(lst),
                             "reference_s": 2.59, "note": "device work + building the Python Particle list"}
print(json.dumps(out))
