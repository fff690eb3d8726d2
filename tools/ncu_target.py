#!/usr/bin/env python
"""Profiling target: ONE batch of BASELINE config 2 (1e5 x 10 GeV photons into lead) with an explicit stack capacity, so that
no pilot/calibration run precedes it and the k-th launch of every wave kernel is wave k+1 of the step:

    ncu --set full --import-source on --clock-control none -k regex:k_loop -s 20 -c 1 -o out python tools/ncu_target.py

captures the plateau wave (wave 21, ~5.6 million records).  ``--dark`` runs config 3 instead (10 GeV e- into graphite,
mV = 3 MeV, 2e4 primaries) and brackets ONLY its dark pass with cudaProfilerStart/Stop: add ``--profile-from-start off``
to the ncu command to capture the dark-pass kernels (k_dark_prepare, k_bucket_*, k_sample, k_dark_emit)."""
import argparse, os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from petite_b200.shower import Shower
from petite_b200.constants import m_electron

DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "data", "")
ap = argparse.ArgumentParser()
ap.add_argument("--primaries", type=int, default=100_000)
ap.add_argument("--dark", action="store_true")
ap.add_argument("--profiling", type=int, default=0, help="2 with PB_LOG_WAVES=<file>: every wave's sizes and per-kernel times")
a = ap.parse_args()
n = a.primaries
if a.dark:
    from petite_b200.dark_shower import DarkShower
    n = min(n, 20_000)
    sh = DarkShower(DATA, "graphite", 0.010, 0.003, seed=20261017, active_processes=["DarkBrem", "DarkAnn", "DarkComp"])
    E0, pid, mass, per = 10.0, 11, m_electron, 900
else:
    sh = Shower(DATA, "lead", 0.010, seed=20261017)
    E0, pid, mass, per = 10.0, 22, 0.0, 1800
if a.profiling:
    sh.set_profiling(a.profiling)
p = np.tile([E0, 0.0, 0.0, np.sqrt(E0 ** 2 - mass ** 2)], (n, 1))
b = sh.run_arrays(p, np.zeros((n, 3)), np.ones(n), np.full(n, mass), np.full(n, pid, dtype=np.int32), np.zeros(n, dtype=np.int32),
                  capacity=n * per, first_shower_id=0)
print("records", b.n, b.counters)
if a.dark:
    import torch
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    dk = sh.generate_dark_showers(b)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print("dark vectors", dk.n, dk.counters)
