#!/usr/bin/env python
"""Per-tile timeline of k_sample in chosen waves of BASELINE config 2 (PB_TILE_LOG, engine.cu): where does the tail of the
persistent sampling kernel come from?  For every wave: span of the launch, busy share of the CTA slots, the longest tiles
(process, map row, samples, trials, microseconds) and what the tiles that END in the last quarter of the span are.

    python tools/tile_timeline.py [primaries] [wave ...] > gpurun_out/tile_timeline.txt"""
import os, sys, tempfile
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "data", "")
N = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
WAVES = [int(w) for w in sys.argv[2:]] or [6, 12, 22, 40]
NAMES = ["Brem", "Ann", "PairProd", "Comp", "Moller", "Bhabha", "MuonE", "MuonBrem"]
from petite_b200.shower import Shower
p = np.tile([10.0, 0.0, 0.0, 10.0], (N, 1))
args = [p, np.zeros((N, 3)), np.ones(N), np.zeros(N), np.full(N, 22, np.int32), np.zeros(N, np.int32)]
for wave in WAVES:
    path = tempfile.mktemp(suffix=".bin")
    os.environ["PB_TILE_LOG"] = f"{path}:{wave}"
    sh = Shower(DATA, "lead", 0.010, seed=20261017)
    sh.run_arrays(*args, capacity=N * 1800, first_shower_id=0)          # warm-up: scratch growth
    os.remove(path) if os.path.exists(path) else None
    sh.run_arrays(*args, capacity=N * 1800, first_shower_id=0)
    raw = np.fromfile(path, dtype=np.uint64).reshape(-1, 4)
    raw = raw[-(1 << 17):]
    ok = raw[:, 1] > 0
    t0, t1 = raw[ok, 0].astype(np.int64), raw[ok, 1].astype(np.int64)
    bucket, count = (raw[ok, 2] >> np.uint64(32)).astype(np.int64), (raw[ok, 2] & np.uint64(0xffffffff)).astype(np.int64)
    trials, sm = (raw[ok, 3] >> np.uint64(16)).astype(np.int64), (raw[ok, 3] & np.uint64(0xffff)).astype(np.int64)
    proc, row = bucket // 256, bucket % 256
    start, end = t0.min(), t1.max()
    span = (end - start) / 1e3
    dur = (t1 - t0) / 1e3
    print(f"== wave {wave}: {ok.sum()} tiles, {count.sum()} samples, {trials.sum()} trials, span {span:.1f} us, sum of tile times {dur.sum():.0f} us "
          f"= {dur.sum() / span:.1f} CTA slots busy on average; median tile {np.median(dur):.1f} us, p99 {np.percentile(dur, 99):.1f}, max {dur.max():.1f}")
    last_sm = np.array([t1[sm == k].max() for k in np.unique(sm)])
    q = (last_sm - start) / 1e3
    print(f"   per-SM finish time: min {q.min():.1f} median {np.median(q):.1f} p90 {np.percentile(q, 90):.1f} max {q.max():.1f} us")
    for k in np.argsort(-dur)[:8]:
        print(f"   long tile: {NAMES[proc[k]]:9s} row {row[k]:3d} samples {count[k]:3d} trials {trials[k]:6d} ({trials[k] / max(count[k], 1):6.1f}/sample) "
              f"{dur[k]:7.1f} us  starts at {(t0[k] - start) / 1e3:7.1f}  ends at {(t1[k] - start) / 1e3:7.1f}")
    late = (t1 - start) / 1e3 > 0.75 * span
    print(f"   tiles ending in the last quarter: {late.sum()}; by process:", {NAMES[q_]: int((late & (proc == q_)).sum()) for q_ in np.unique(proc[late])},
          f"mean duration {dur[late].mean():.1f} us, mean samples {count[late].mean():.0f}")
    for q_ in np.unique(proc):
        s = proc == q_
        print(f"   {NAMES[q_]:9s} tiles {s.sum():6d} us/tile {dur[s].mean():7.1f} us/trial-per-CTA {1e3 * dur[s].sum() / max(trials[s].sum(), 1):7.2f} ns trials/sample {trials[s].sum() / count[s].sum():6.1f}")
    del sh
