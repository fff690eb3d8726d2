"""Tuning + safety sweep of the accept/reject sampler (one process, one engine per configuration; the knobs are read from
the environment by pb_create).  A configuration is "G,T,generic": G = lanes per sample (PB_SAMPLE_G), T = trials per lane
and round (PB_SAMPLE_T), generic = 1 runs the SM pass through the kernel that also carries the dark integrands
(PB_SAMPLE_GENERIC).  For every configuration

  1. correctness: 2 000 showers of config 2 must give the SAME records and trial counts as the first configuration - the
     draws are counter-based, so any schedule has to reproduce them bit for bit;
  2. timing: config 2 at 1e5 primaries, CUDA events, per-step and per-kernel (SWEEP_PROFILING=1: the two loop kernels,
     2: every kernel).

    python tools/sweep_sampler.py [n_timing] [G,T,generic ...] > gpurun_out/sweep_sampler.log
    PETITE_B200_LIB=variants/libpb_X.so python tools/sweep_sampler.py ...      # compile-time variants: tools/sweep_variants.sh
"""
import os, sys, json, hashlib
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from petite_b200.shower import Shower

DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "data", "")
N_T = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
CONFIGS = [(4, 1, 1), (4, 1, 0), (2, 2, 0), (1, 2, 0), (4, 2, 0)]
if len(sys.argv) > 2:
    CONFIGS = [tuple(int(v) for v in c.split(",")) for c in sys.argv[2:]]


def prim(k, dev=None):
    p = np.tile(np.array([10.0, 0, 0, 10.0]), (k, 1)); r = np.zeros((k, 3)); w = np.ones(k); m = np.zeros(k)
    a = [p, r, w, m, np.full(k, 22, np.int32), np.zeros(k, np.int32)]
    if dev is not None:
        a = [torch.from_numpy(x).to(dev) for x in a]
    return a


def digest(sh):
    b = sh.run_arrays(*prim(2000), first_shower_id=777_000)
    h = b.to_host()
    o = np.lexsort((h["p0"][:, 3], h["p0"][:, 0], h["generation"], h["shower"]))
    return dict(b.counters), h["p0"][o].copy(), h["pf"][o].copy(), h["ntrials"][o].copy(), h["pid"][o].copy()


dev = torch.device("cuda", 0)
ref = None
for G, T, generic in CONFIGS:
    os.environ["PB_SAMPLE_G"], os.environ["PB_SAMPLE_T"], os.environ["PB_SAMPLE_GENERIC"] = str(G), str(T), str(generic)
    sh = Shower(DATA, "lead", 0.010, seed=20261017)
    d = digest(sh)
    if ref is None:
        ref = d
    same = all(d[0][k] == ref[0][k] for k in ("n_particles", "n_steps", "n_substeps", "n_samples", "n_trials", "n_no_sample")) and \
        all(np.array_equal(a, b) for a, b in zip(d[1:], ref[1:]))
    devp = prim(N_T, dev)
    cal = sh.run_arrays(*prim(2000), first_shower_id=10 ** 9)
    cap = int(N_T * cal.n / 2000 * 1.06 + 2.3 * cal.counters["max_wave"] / 2000 * N_T) + (1 << 16)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sh.run_arrays(*devp, capacity=cap, first_shower_id=0)          # warm-up (scratch growth)
    sh.set_profiling(int(os.environ.get("SWEEP_PROFILING", "1")))
    ms, ks, kl = [], [], []
    for rep in range(3):
        torch.cuda.synchronize(); e0.record()
        b = sh.run_arrays(*devp, capacity=cap, first_shower_id=0)
        e1.record(); torch.cuda.synchronize()
        pr = sh.get_profile()
        ms.append(e0.elapsed_time(e1)); ks.append(pr["ms"]["k_sample"]); kl.append(pr["ms"]["k_loop"])
    sh.set_profiling(0)                                            # the wave loop as one CUDA graph (the bench's main region)
    mg = []
    for rep in range(3):
        torch.cuda.synchronize(); e0.record()
        b = sh.run_arrays(*devp, capacity=cap, first_shower_id=0)
        e1.record(); torch.cuda.synchronize()
        mg.append(e0.elapsed_time(e1))
    sha = hashlib.sha1(b"".join(np.ascontiguousarray(a).tobytes() for a in d[1:])).hexdigest()[:12]
    print(json.dumps({"drain": os.environ.get("PB_DRAIN_LANES", ""), "graph_step_ms": round(min(mg), 2), "waves": b.counters.get("n_waves"), "digest": sha, "G": G, "T": T, "generic": generic, "lib": os.path.basename(os.environ.get("PETITE_B200_LIB", "default")), "same_as_reference": bool(same), "records": d[0]["n_particles"],
                      "trials": d[0]["n_trials"], "step_ms": round(min(ms), 2), "k_sample_ms": round(min(ks), 2),
                      "k_loop_ms": round(min(kl), 2), "other": {k: round(v, 2) for k, v in pr["ms"].items() if v and k not in ("k_sample", "k_loop")}, "showers_per_s": round(N_T / min(ms) * 1e3)}), flush=True)
    del sh, b, cal, devp
    torch.cuda.empty_cache()
