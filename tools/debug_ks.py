import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from scipy.stats import ks_2samp
from tests.gpu_util import primaries
from tests.test_gpu_dark import dark_shower, oracle_dark
ds = dark_shower("graphite", 0.03)
n_gpu, n_orc = 3000, 60
prims = primaries(11, 2.0, n_gpu)
sm = ds.generate_showers(prims, first_shower_id=50_000)
dk = ds.generate_dark_showers(sm)
h = dk.to_host()
ref = oracle_dark(prims[:n_orc], "graphite", 0.03, 0.010, 31, 0)
y_gpu = np.bincount(h["shower"], weights=h["weight"], minlength=n_gpu)
y_orc = np.array([sum(v.weight for v in vs) for _, vs in ref])
print("yield", ks_2samp(y_gpu, y_orc).pvalue)
print("count", ks_2samp(np.bincount(h["shower"], minlength=n_gpu), np.array([len(vs) for _, vs in ref])).pvalue)
w_orc = np.array([v.weight for _, vs in ref for v in vs])
E_orc = np.array([v.p0[0] for _, vs in ref for v in vs])
print("full w", ks_2samp(h["weight"], w_orc).pvalue, "full E", ks_2samp(h["p0"][:, 0], E_orc).pvalue, len(w_orc))
st = max(1, dk.n // 20000)
for off in range(0, st, max(1, st // 6)):
    print(" stride off", off, ks_2samp(h["weight"][off::st], w_orc).pvalue, ks_2samp(h["p0"][off::st, 0], E_orc).pvalue)
