"""Debug aid: run the same SM + dark batch several times on one engine (with other batches in between) and check that the
results are bit-identical (records, weights), as the counter-based RNG promises."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from tests.gpu_util import primaries
from tests.test_gpu_dark import dark_shower


def run(ds, n, E, first):
    prims = primaries(11, E, n)
    sm = ds.generate_showers(prims, first_shower_id=first)
    dk = ds.generate_dark_showers(sm)
    h = dk.to_host()
    hs = sm.to_host()
    o = np.lexsort((hs["p0"][:, 0], hs["shower"]))
    od = np.lexsort((h["p0"][:, 0], h["weight"], h["shower"]))
    return sm.n, dk.n, hs["p0"][o].copy(), h["weight"][od].copy(), h["p0"][od].copy(), dict(sm.counters)


ds = dark_shower("graphite", 0.03)
ref = None
for it in range(4):
    if it == 1:
        run(ds, 3, 5.0, 7000)           # a small batch in between, as the test suite does
    if it == 2:
        run(ds, 20000, 1.0, 9000)
    r = run(ds, 3000, 2.0, 50000)
    print(it, r[0], r[1], r[5])
    if ref is None:
        ref = r
    else:
        print("  same counts", r[0] == ref[0], r[1] == ref[1], "sm p0 equal", r[0] == ref[0] and np.array_equal(r[2], ref[2]),
              "dark w equal", r[1] == ref[1] and np.array_equal(r[3], ref[3]))
