"""Headline benchmark: showers/s (and particle-steps/s) of the shower-stepping hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config 1..5] [--scaling weak|strong]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path over one batch of synthetic primaries.  Default workload: BASELINE.json configs[1], 10 GeV
photons into lead, E_min = 10 MeV, 1e5 primaries per GPU (weak scaling: every rank steps its own 1e5 showers, shower ids offset
by rank; ``--scaling strong``: the 1e5 primaries are split over the ranks).  A step is one batch on one stream, the wave loop one CUDA
graph launch (``--parts 2`` steps it as two concurrent sub-batches: better only for the narrow-wave configurations 1 and 5 since the
kernel tails were removed in round 2).  The only collective is one NCCL all-reduce of the
8 KB tally buffer per step.  ``--config`` selects another BASELINE.json configuration (1, 3, 4, 5; the dark configurations run
their dark pass inside the step and are stepped in sub-batches that fit HBM, tallies only).

One JSON line on rank 0:  value = whole-job showers/s with primaries resident in HBM; e2e = the same through the public
host API (pinned host primaries copied in, tallies read back); e2e_history = a step that also copies the WHOLE particle
history back to the host (what the reference's generate_shower returns); maxF_fudge_4 = the same workload at the reference's
observed acceptance rate; retrained_maps = the same workload after Shower.retrain_maps() (maps trained on this GPU for
the sampler's own figure of merit: same physics, fewer trials); roofline = the dominant kernel against the measured HBM peak as the contract asks,
and "fp64" = the same kernel against a measured FP64 FMA peak (the path is FP64-pipe bound), with flops COUNTED
in the oracle (oracle/count_ops.py -> petite_b200/roofline.py); cpu_baseline = the CPU oracle (a port of the reference's
generate_shower) on the host cores over a bounded sample.  ``--impl reference`` times that CPU path alone.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
DATA = os.path.join(ROOT, "data", "")
DATA400 = os.path.join(ROOT, "data_400GeV", "")
SEED, EMIN = 20261017, 0.010
M_E, M_MU = 510.998950e-6, 105.6583755e-3

# BASELINE.json configs (SURVEY.md 8d gives them as concrete synthetic inputs).  batch = primaries stepped per engine call
# (the 100 GeV / 400 GeV showers keep 6-8 thousand records each, so 1e5 of them do not fit HBM at once).
CONFIGS = {
    1: dict(workload="SM shower: 10 GeV e- into graphite, E_min=0.010 GeV, ./data/ maps, 1e3 primaries (BASELINE.json configs[0])",
            material="graphite", pid=11, E0=10.0, mass=M_E, mV=None, n=1_000, batch=1_000, cpu=8, parts=2),
    2: dict(workload="SM shower: 10 GeV photon into lead, E_min=0.010 GeV, 1e5 primaries (BASELINE.json configs[1])",
            material="lead", pid=22, E0=10.0, mass=0.0, mV=None, n=100_000, batch=100_000, cpu=8),
    3: dict(workload="Dark shower: 10 GeV e- into graphite, m_V=3 MeV (lightest trained mass; the literal 1 MeV runs as 1 GeV, SURVEY Q-2), "
                     "DarkBrem/DarkAnn/DarkComp, 1e5 primaries (BASELINE.json configs[2])",
            material="graphite", pid=11, E0=10.0, mass=M_E, mV=0.003, active=["DarkBrem", "DarkAnn", "DarkComp"], n=100_000, batch=100_000, cpu=2),
    4: dict(workload="High-energy dark shower: data_400GeV maps, synthetic beam-dump e+-/gamma spectrum into lead, m_V=10 MeV, 1e5 primaries "
                     "(BASELINE.json configs[3])",
            material="lead", pid=0, E0=0.0, mass=0.0, mV=0.010, active=["DarkBrem", "DarkAnn", "DarkComp"], n=100_000, batch=20_000, data=DATA400, cpu=1),
    5: dict(workload="Muon dark shower: 100 GeV mu- through lead, MuonBrem/MuonE/DarkMuonBrem + multiple scattering, m_V=30 MeV, 1e5 primaries "
                     "(BASELINE.json configs[4])",
            material="lead", pid=13, E0=100.0, mass=M_MU, mV=0.030, active=["DarkMuonBrem", "DarkBrem", "DarkAnn", "DarkComp"], n=100_000, batch=25_000, cpu=1, parts=2),
}
CFG = CONFIGS[2]


def beam_dump_spectrum(n, seed=SEED):
    """Config 4: 50 % photons resampled from the reference's 120 GeV pi0-photon beam scaled x(400/120) in
    momentum, 25 % e-, 25 % e+ with dN/dE ~ 1/E on [1, 400] GeV along +z."""
    rng = np.random.default_rng(seed)
    beam = np.load(DATA400 + "Photons_From_Pi0s_120GeV.npy")
    g = beam[rng.integers(0, len(beam), n // 2)] * (400.0 / 120.0)
    g = g[g[:, 0] > 0.0016]
    ne = n - len(g)
    E = np.exp(rng.uniform(np.log(1.0), np.log(400.0), ne))
    pe = np.column_stack([E, np.zeros(ne), np.zeros(ne), np.sqrt(E ** 2 - M_E ** 2)])
    pid = np.concatenate([np.full(len(g), 22), np.where(np.arange(ne) % 2 == 0, 11, -11)]).astype(np.int32)
    return np.vstack([g, pe]), pid, np.where(pid == 22, 0.0, M_E)


def make_primaries(cfg, n):
    """-> (p (n,4), r (n,3), w, m, pid, flags) NumPy arrays of the configuration's synthetic beam."""
    if cfg["pid"] == 0:
        p, pid, m = beam_dump_spectrum(n)
    else:
        pz = np.sqrt(cfg["E0"] ** 2 - cfg["mass"] ** 2)
        p = np.tile([cfg["E0"], 0.0, 0.0, pz], (n, 1)); pid = np.full(n, cfg["pid"], np.int32); m = np.full(n, cfg["mass"])
    return p, np.zeros((n, 3)), np.ones(n), m, pid, np.zeros(n, np.int32)


# ------------------------------------------------------------------------------------------------ CPU arm
_ORC = None
_CPU_PRIM = None


def _cpu_init(cfg_id):
    """Worker set-up: the oracle (SM or dark) and the configuration's primaries (config 4 draws a spectrum)."""
    global _ORC, _CPU_PRIM, CFG
    CFG = CONFIGS[cfg_id]
    data = CFG.get("data")
    if CFG["mV"] is None:
        from oracle.shower import OracleShower
        _ORC = OracleShower(data, CFG["material"], EMIN, seed=SEED, rng="counter")
    else:
        from oracle.dark import OracleDarkShower
        _ORC = OracleDarkShower(data, CFG["material"], EMIN, CFG["mV"], active_processes=CFG["active"], seed=SEED, rng="counter")
    _CPU_PRIM = make_primaries(CFG, 4096 if CFG["pid"] == 0 else 1)


def _cpu_one(i):
    from oracle.shower import OParticle
    p, _, _, m, pid, _ = _CPU_PRIM
    k = i % len(pid)
    sh = _ORC.generate_shower(OParticle(list(p[k]), [0, 0, 0], PID=int(pid[k]), ID=1, mass=float(m[k])), shower_id=i)
    steps = sum(1 for q in sh if q.ended)
    n_dark = 0
    if CFG["mV"] is not None:
        n_dark = len(_ORC.generate_dark_shower(sh)[1])
    return len(sh), steps, n_dark


def cpu_run(cfg_id, n_showers, cores, first_id=0, pool=None):
    """Oracle showers over a process pool -> (seconds, particles, steps)."""
    import multiprocessing as mp
    own = pool is None
    if own:
        pool = mp.get_context("fork").Pool(cores, initializer=_cpu_init, initargs=(cfg_id,))
        pool.map(_cpu_one, range(cores))            # construct the tables in every worker before timing
    t0 = time.perf_counter()
    res = pool.map(_cpu_one, range(first_id, first_id + n_showers), chunksize=1)
    dt = time.perf_counter() - t0
    if own:
        pool.close()
    return dt, sum(r[0] for r in res), sum(r[1] for r in res)


def config_block(cfg, extra):
    d = {"workload": cfg["workload"], "material": cfg["material"], "pid": cfg["pid"] or "spectrum (50% gamma, 25% e-, 25% e+)",
         "E0_GeV": cfg["E0"] or "1-400 (beam-dump spectrum)", "E_min_GeV": EMIN}
    if cfg["mV"] is not None:
        d["mV_GeV"] = cfg["mV"]; d["dark_processes"] = cfg["active"]
    d.update(extra)
    return d


def reference_arm(args):
    """--impl reference: the reference's CPU algorithm (oracle port; the reference itself cannot be installed here,
    see DESIGN.md) on all host cores, same metric and config; each step is a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    cfg = CONFIGS[args.config]
    cores = os.cpu_count() or 1
    pool = mp.get_context("fork").Pool(cores, initializer=_cpu_init, initargs=(args.config,))
    pool.map(_cpu_one, range(cores))
    n = args.cpu_showers or max(cfg["cpu"] * cores, 8 * cfg["cpu"])
    for w in range(args.warmup):
        cpu_run(args.config, n, cores, first_id=10_000 + w * n, pool=pool)
    tot_t, tot_p, tot_s = 0.0, 0, 0
    for k in range(args.steps):
        dt, npart, nsteps = cpu_run(args.config, n, cores, first_id=k * n, pool=pool)
        tot_t += dt; tot_p += npart; tot_s += nsteps
    pool.close()
    v = args.steps * n / tot_t
    line = {"impl": "reference", "metric": "showers/sec", "value": v, "unit": "showers/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / args.steps, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_block(cfg, {"showers_per_step": n, "note": "CPU oracle port of PETITE generate_shower / generate_dark_shower "
                                                                          "(vectorised sweeps), multiprocessing over all host cores"}),
            "particle_steps_per_sec": tot_s / tot_t,
            "cpu_baseline": {"value": v, "unit": "showers/s", "cores": cores, "kind": "port",
                             "sample": f"{n} showers/step x {args.steps} steps of the same workload"},
            "e2e": {"value": v, "unit": "showers/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ GPU arm
def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ours(args):
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cfg = CONFIGS[args.config]
    dark = cfg["mV"] is not None

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        # before CUDA is initialised in this process (fork-safe)
        cores = os.cpu_count() or 1
        n_cpu = args.cpu_showers or max(cfg["cpu"] * cores, 8 * cfg["cpu"])
        dt, npart, nsteps = cpu_run(args.config, n_cpu, cores)
        cpu = {"value": n_cpu / dt, "unit": "showers/s", "cores": cores, "kind": "port",
               "sample": f"{n_cpu} showers of the same workload through the CPU oracle (multiprocessing, {cores} procs), {dt:.1f} s",
               "particle_steps_per_sec": nsteps / dt}

    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from petite_b200.shower import Shower
    from petite_b200 import roofline as rl
    from petite_b200 import _capi as capi
    from petite_b200.distributed import shard

    dev = torch.device("cuda", local_rank)
    data = cfg.get("data", DATA)

    def engine(fudge=1):
        if not dark:
            return Shower(data, cfg["material"], EMIN, maxF_fudge_global=fudge, seed=SEED, device=local_rank)
        from petite_b200.dark_shower import DarkShower
        return DarkShower(data, cfg["material"], EMIN, cfg["mV"], maxF_fudge_global=fudge, active_processes=cfg["active"], seed=SEED,
                          device=local_rank)

    sh = engine()
    # ---- this rank's primaries.  weak: n per GPU (global showers [rank n, (rank + 1) n)); strong: the configuration's n split over the ranks
    n_cfg = args.primaries or cfg["n"]
    if args.scaling == "strong":
        base_id, n = shard(n_cfg, rank, world)
        n_job = n_cfg
    else:
        n, n_job = n_cfg, world * n_cfg
        base_id = rank * n
    allp = make_primaries(cfg, n_job if cfg["pid"] == 0 else n)
    if cfg["pid"] == 0:
        allp = tuple(a[base_id:base_id + n] for a in allp)
        n = len(allp[4])
    host = [torch.from_numpy(np.ascontiguousarray(a)).pin_memory() for a in allp]
    host_np = [t.numpy() for t in host]
    devp = [t.to(dev) for t in host]
    h2d_bytes = sum(t.numel() * t.element_size() for t in host)
    batch = max(1, min(cfg["batch"], n))
    slices = [slice(a, min(a + batch, n)) for a in range(0, n, batch)]

    # ---- calibration: records and widest wave per GeV of primary energy -> stack capacity of the largest sub-batch
    ncal = min(n, 2000 if cfg["E0"] <= 10.0 else 256)
    idx = np.unique(np.linspace(0, n - 1, ncal).astype(np.int64))
    cal = sh.run_arrays(*[a[idx] for a in host_np], first_shower_id=10 ** 9)
    e_cal = float(host_np[0][idx, 0].sum())
    e_max = max(float(host_np[0][sl, 0].sum()) for sl in slices)
    slack = 1.06 if cfg["pid"] != 0 else 1.20
    capacity = int(e_max / e_cal * (cal.n * slack + 2.3 * cal.counters["max_wave"])) + (1 << 16)
    del cal
    sh._ensure_stack(capacity)
    tally = torch.zeros(capi.TALLY_SIZE, dtype=torch.float64, device=dev)
    dtally = torch.zeros(capi.TALLY_SIZE, dtype=torch.float64, device=dev) if dark else None

    P = args.parts if args.parts > 0 else cfg.get("parts", 1)
    COUNT_KEYS = ("n_particles", "n_steps", "n_substeps", "n_samples", "n_trials", "n_launches", "n_waves", "n_charged")

    def step(arrays, parts=1, eng=None, keep=False):
        """One pass over this rank's primaries: every sub-batch through the SM wave loop (+ the dark pass), tallies accumulated on
        the device, one all-reduce.  -> summed counters (+ the last batches if ``keep``)."""
        eng = eng or sh
        tally.zero_()
        if dark:
            dtally.zero_()
        tot = {k: 0 for k in COUNT_KEYS}
        tot.update(n_dark=0, n_dark_trials=0, n_dark_samples=0, max_wave=0)
        kept = []
        for sl in slices:
            part = [a[sl] for a in arrays]
            first = base_id + sl.start
            if parts > 1:                 # concurrent sub-batches: one engine handle, stream and host thread per part, tallied per part
                bs = eng.run_arrays_split(*part, parts=parts, capacity=capacity, first_shower_id=first, tally=tally)
            else:
                bs = [eng.run_arrays(*part, capacity=capacity, first_shower_id=first)]
                eng.tally_batches(bs, tally)
            tot["n_launches"] += len(bs)                                   # k_tally per part
            for b in bs:
                for key in COUNT_KEYS:
                    tot[key] += b.counters[key]
                tot["max_wave"] = max(tot["max_wave"], b.counters["max_wave"])
                if dark:
                    d = eng.generate_dark_showers(b)
                    eng.tally_dark(d, dtally)
                    tot["n_dark"] += d.n; tot["n_dark_trials"] += d.counters["n_trials"]; tot["n_dark_samples"] += d.counters["n_samples"]
                    tot["n_launches"] += d.counters["n_launches"] + 1
                    del d                                                  # (a live DarkBatch pins its dark stack)
            if keep:
                kept = bs
        if world > 1:
            dist.all_reduce(tally)        # the only collective: 8 KB of tallies over NVLink
            if dark:
                dist.all_reduce(dtally)
            tot["n_launches"] += 2 if dark else 1
        return (tot, kept) if keep else tot

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # warm-up (single stream; in split mode also the peer engines: tables, stacks, scratch growth); the last single-stream
    # warm-up step times EVERY kernel (profiling level 2, ~6 % overhead) for the per-kernel table (SM pass of the last sub-batch)
    full_ms, full_launch = {}, {}
    for w in range(args.warmup):
        if w == args.warmup - 1 and not dark:
            sh.set_profiling(2)
        step(devp)
    if args.warmup and not dark:
        pr = sh.get_profile()
        scale = len(slices)
        full_ms, full_launch = {k: v * scale for k, v in pr["ms"].items()}, {k: v * scale for k, v in pr["launches"].items()}
    sh.set_profiling(0)
    if P > 1:
        for w in range(max(args.warmup, 1)):
            step(devp, P)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timed_region(parts, level, steps):
        """``steps`` steps bracketed by barrier + synchronize; -> (ms max over ranks, summed counters, per-kernel ms / launches, trials)."""
        level = 0 if dark or len(slices) > 1 else level      # the engine keeps one profile per call: per-kernel times need one call per step
        sh.set_profiling(level)
        pm = {k: 0.0 for k in capi.KERNEL_NAMES}
        pl = {k: 0 for k in capi.KERNEL_NAMES}
        tr = {}
        tt = {}
        barrier()
        e0.record()
        for k in range(steps):
            c = step(devp, parts)
            if level:
                pr = sh.get_profile()
                for name in pm:
                    pm[name] += pr["ms"][name]; pl[name] += pr["launches"][name]
                for p_, v in pr["trials"].items():
                    tr[p_] = tr.get(p_, 0) + v
            for key, v in c.items():
                tt[key] = max(tt.get(key, 0), v) if key == "max_wave" else tt.get(key, 0) + v
        e1.record()
        barrier()
        t_ms = e0.elapsed_time(e1)
        sh.set_profiling(0)
        if world > 1:
            t = torch.tensor([t_ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            t_ms = float(t.item())
        return t_ms, tt, pm, pl, tr

    # ---- timed region: K steps, device-resident primaries.  With --parts P > 1 (the default of configurations 1 and 5) a step runs as P concurrent
    # sub-batches and carries no per-kernel events; the single-stream pass after it (same K steps, same inputs) times the two
    # dominant kernels with CUDA events (level 1) while each launch owns the GPU: that is what the roofline is defined on.
    K = args.steps
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    ms, tot, _, _, _ = timed_region(P, 0, K)              # no per-kernel events: the wave loop runs as one CUDA graph per batch
    clk = clocks.stop() if rank == 0 else None
    # the same K steps as ONE batch on one stream with CUDA events around the two dominant kernels (profiling level 1 switches the
    # engine to stream launches, one host synchronisation per growing wave): what the per-kernel, roofline and fp64 figures come from
    ms1, tot1, prof_ms, prof_launch, trials = timed_region(1, 1, K)
    single = {"value": n_job * K / (ms1 * 1e-3), "unit": "showers/s", "ms_per_step": ms1 / K,
              "note": "the same K steps as ONE batch on one stream, stream launches with per-kernel CUDA events: the region the per-kernel times, "
                      "roofline and fp64 figures come from"}
    launches_main = tot["n_launches"]
    tot = dict(tot1, n_launches=launches_main)          # identical showers: only the launch count differs
    tally_host = tally.cpu().numpy()
    dtally_host = dtally.cpu().numpy() if dark else None

    # ---- end-to-end through the public host API: pinned host primaries in, tallies out, every step
    e2e_steps = max(1, min(K, 3))
    step(host_np, P)            # untimed: host-staging buffers of every engine handle at full size
    barrier()
    e0.record()
    for k in range(e2e_steps):
        step(host_np, P)
        _ = tally.cpu()
        if dark:
            _ = dtally.cpu()
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms_e2e], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_e2e = float(t.item())

    # ---- end to end WITH the particle history: what the reference's generate_shower returns is every particle, so this step also
    # streams the filled part of the stack (p0, r0w, pf, rf, ids, aux: 168 B per record) to the host through two pinned staging
    # buffers.  One step, single batch per engine call, rank-local (no collective).
    hist = None
    if args.history and rank == 0 and world == 1:            # the three extra lines below are single-GPU measurements (no rank may wait on rank 0)
        CH = 1 << 21                                                  # records per chunk (2 Mi x 32 B = 64 MiB per column chunk)
        stage = [torch.empty(CH * 32, dtype=torch.uint8, pin_memory=True) for _ in range(2)]
        evs = [torch.cuda.Event(), torch.cuda.Event()]
        copy_stream = torch.cuda.Stream(device=dev)

        def drain(b):
            nbytes, k = 0, 0
            for name in ("p0", "r0w", "pf", "rf", "ids", "aux"):
                col = b._t[name][:b.n]
                flat = col.view(torch.uint8).reshape(-1)
                row = col.shape[1] * col.element_size()
                for a in range(0, b.n, CH):
                    m_ = min(CH, b.n - a)
                    evs[k % 2].synchronize()                        # the staging buffer's previous copy has landed
                    stage[k % 2][: m_ * row].copy_(flat[a * row:(a + m_) * row], non_blocking=True)
                    evs[k % 2].record()
                    nbytes += m_ * row; k += 1
            return nbytes
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        hbytes = 0
        tally.zero_()
        for sl in slices:
            b = sh.run_arrays(*[a[sl] for a in host_np], capacity=capacity, first_shower_id=base_id + sl.start)
            sh.tally(b, tally)
            hbytes += drain(b)
        _ = tally.cpu()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        hist = {"value": n / dt, "unit": "showers/s", "d2h_bytes_per_step": int(hbytes), "h2d_bytes_per_step": h2d_bytes, "seconds": dt,
                "d2h_GBps": hbytes / dt / 1e9, "note": "one GPU, one step, SM history only: run_arrays(host arrays) + every stack column of the "
                "filled records copied to pinned host memory (2 x 64 MiB staging), wall clock"}

    # ---- the same workload at the reference's observed acceptance rate: maxF_fudge_global = 4 (the authors' tables
    # give ~90 trials per sample where the regenerated max_F gives ~15; both arms of every ratio here use the regenerated tables)
    fudge4 = None
    if args.fudge_line and rank == 0 and world == 1:
        sh4 = engine(4)
        sh4._stack_tensors, sh4._stack_capacity = sh._stack_tensors, sh._stack_capacity       # same HBM: the runs do not overlap
        if dark:
            sh4._dark_stack, sh4._dark_capacity = sh._dark_stack, sh._dark_capacity
        step(devp, 1, eng=sh4)
        torch.cuda.synchronize()
        e0.record()
        c4 = {}
        for k in range(2):
            c = step(devp, 1, eng=sh4)
            for key, v in c.items():
                c4[key] = c4.get(key, 0) + v
        e1.record()
        torch.cuda.synchronize()
        ms4 = e0.elapsed_time(e1) / 2
        fudge4 = {"value": n / (ms4 * 1e-3), "unit": "showers/s", "ms_per_step": ms4, "trials_per_sample": c4["n_trials"] / max(c4["n_samples"], 1),
                  "note": "maxF_fudge_global=4, one GPU, one stream, device-resident primaries (compare single_stream)"}
        del sh4

    # ---- the same workload on maps retrained HERE (rows f-2 + f-1: Shower.retrain_maps, training weight |jac f|^8): what a user gets
    # after one call that takes a few seconds on the GPU.  Not the headline - value / e2e use the reference's shipped maps.
    retr = None
    if args.retrained_line and rank == 0 and world == 1 and not dark:
        shr = engine()
        shr._stack_tensors, shr._stack_capacity = sh._stack_tensors, sh._stack_capacity
        t0 = time.perf_counter()
        gain = shr.retrain_maps()
        t_train = time.perf_counter() - t0
        step(devp, 1, eng=shr)
        torch.cuda.synchronize()
        e0.record()
        cr = {}
        for k in range(2):
            c = step(devp, 1, eng=shr)
            for key, v in c.items():
                cr[key] = cr.get(key, 0) + v
        e1.record()
        torch.cuda.synchronize()
        msr = e0.elapsed_time(e1) / 2
        retr = {"value": n / (msr * 1e-3), "unit": "showers/s", "ms_per_step": msr, "trials_per_sample": cr["n_trials"] / max(cr["n_samples"], 1),
                "records_per_shower": cr["n_particles"] / (2 * n), "retrain_seconds": t_train,
                "accept_rate_over_shipped_maps": {P: float(np.exp(np.nanmean(np.log(g[1][np.isfinite(g[1]) & (g[1] > 0)])))) for P, g in gain.items()},
                "note": "Brem and PairProd maps retrained on this GPU (Shower.retrain_maps: VEGAS refinement on |jac f|^8, then find_max), one "
                        "stream, device-resident primaries (compare single_stream / value); same physics, fewer accept/reject trials"}
        del shr

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    hbm_peak, peak_src = measured_peaks()
    fp64_peak = sh.measure_fp64_peak()
    tot["n_daughters"] = tot["n_particles"] - K * n
    for k in prof_ms:                                           # kernels not timed live: scale the warm-up measurement
        if prof_ms[k] == 0.0 and full_ms.get(k):
            prof_ms[k] = full_ms[k] * K
            prof_launch[k] = prof_launch[k] or full_launch[k] * K
    if not trials and not dark:
        trials = {}
    step_ms_total = sum(prof_ms.values())
    dom = max(("k_loop", "k_sample"), key=prof_ms.get)          # dominant kernel of the step (both timed live)
    dom_ms = prof_ms[dom]
    dom_bytes = rl.kernel_bytes(dom, tot)
    ach = dom_bytes / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
    dom_flops = rl.kernel_flops(dom, tot, trials)
    all_bytes = rl.step_bytes(tot["n_particles"], tot["n_charged"], tot["n_samples"], tot["n_daughters"])
    ncu_fp64 = {}
    npath = os.path.join(ROOT, "profiles", "r02_ncu_fp64.json")
    if os.path.exists(npath):
        try:
            ncu_fp64 = json.load(open(npath))
        except Exception:
            ncu_fp64 = {}
    per_kernel = {k: {"ms_per_step": prof_ms[k] / K, "launches_per_step": prof_launch[k] / K,
                      "alg_GBps": (rl.kernel_bytes(k, tot) / (prof_ms[k] * 1e-3) / 1e9) if prof_ms[k] > 0 else 0.0,
                      "model_fp64_tflops": (rl.kernel_flops(k, tot, trials) / (prof_ms[k] * 1e-3) / 1e12) if prof_ms[k] > 0 else 0.0,
                      "ncu_fp64_pipe_pct": ncu_fp64.get(k)}
                  for k in prof_ms}
    traffic = None
    for tname in ("r02_traffic.json", "r01_traffic.json"):
        tpath = os.path.join(ROOT, "profiles", tname)
        if os.path.exists(tpath):
            try:
                traffic = json.load(open(tpath)).get(dom)
            except Exception:
                traffic = None
            if traffic:
                traffic = dict(traffic, source="profiles/" + tname)
                break
    line = {
        "metric": "showers/sec", "value": n_job * K / (ms * 1e-3), "unit": "showers/s", "n_gpus": world, "steps": K,
        "warmup": args.warmup, "ms_per_step": ms / K, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": config_block(cfg, {
            "primaries_per_gpu": n, "primaries_per_step": n_job, "sub_batches_per_step": len(slices),
            "parallelism": f"{world} x independent shower shards, tallies all-reduced" + (f"; each shard stepped as {P} concurrent sub-batches (streams)" if P > 1 else ""),
            "seed": SEED, "maxF": "regenerated (oracle.findmax, B=300)", "wave_loop": "CUDA graph (WHILE node, device-side condition)",
            "l2": f"working set {capacity * rl.RECORD_BYTES / 1e9:.1f} GB of stack per GPU >> 126 MB L2 (no flush needed)",
            "stack_capacity_records": capacity}),
        "particle_steps_per_sec": world * tot["n_steps"] / (ms * 1e-3) if args.scaling == "weak" else None,
        "trials_per_sec": world * tot["n_trials"] / (ms * 1e-3) if args.scaling == "weak" else None,
        "per_shower": {"records": tot["n_particles"] / (K * n), "steps": tot["n_steps"] / (K * n),
                       "substeps": tot["n_substeps"] / (K * n), "trials": tot["n_trials"] / (K * n),
                       "waves": tot["n_waves"] / (K * len(slices))},
        "roofline": {"kernel": dom, "bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s",
                     "frac": ach / hbm_peak, "traffic": traffic["dram_bytes_per_launch"] if traffic else None,
                     "traffic_capture": traffic, "peak_source": peak_src,
                     "launches": prof_launch[dom], "avg_launch_ms": dom_ms / max(prof_launch[dom], 1),
                     "algorithmic_bytes_per_launch": dom_bytes / max(prof_launch[dom], 1),
                     "share_of_step": dom_ms / step_ms_total if step_ms_total else None,
                     "timed_region": "single-stream pass (single_stream): the main region runs the wave loop as a CUDA graph without per-kernel events" + ((", %d sub-batches concurrently" % P) if P > 1 else ""),
                     "note": "the step is FP64-pipe / divergence bound, not HBM bound (SURVEY.md 8d): see fp64"},
        "fp64": {"kernel": dom, "achieved_tflops": dom_flops / (dom_ms * 1e-3) / 1e12 if dom_ms else 0.0,
                 "peak_tflops": fp64_peak, "frac": (dom_flops / (dom_ms * 1e-3) / 1e12 / fp64_peak) if dom_ms and fp64_peak else None,
                 "peak_source": "pb_measure_fp64_peak (DFMA chains, this run)",
                 "specials_per_sec": (rl.sample_kernel_specials(trials) / (dom_ms * 1e-3)) if dom == "k_sample" and dom_ms else None,
                 "ncu_fp64_pipe_pct": ncu_fp64.get(dom),
                 "model": "flops COUNTED in the oracle's restatement of the reference formulas (python -m oracle.count_ops -> petite_b200/roofline.py: "
                          "+,-,*,/,sqrt = 1 flop; cos/sin/exp/log/pow = specials, reported separately) x measured trial / sub-step counters"},
        "whole_step_hbm": {"algorithmic_GBps": all_bytes / (ms * 1e-3) / 1e9, "frac_of_peak": all_bytes / (ms * 1e-3) / 1e9 / hbm_peak},
        "kernels": per_kernel,
        "trials_by_process": trials,
        "e2e": {"value": n_job * e2e_steps / (ms_e2e * 1e-3), "unit": "showers/s", "h2d_bytes_per_step": h2d_bytes,
                "d2h_bytes_per_step": int(tally_host.nbytes * (2 if dark else 1) + 8 * 16), "steps": e2e_steps,
                "api": ("Shower.run_arrays_split(host arrays, parts=%d)" % P if P > 1 else "Shower.run_arrays(host arrays)") + " + pb_tally + tally.cpu()"
                       + (" + DarkShower.generate_dark_showers + tally_dark" if dark else "")},
        "e2e_history": hist,
        "maxF_fudge_4": fudge4,
        "retrained_maps": retr,
        "gpu_launches": tot["n_launches"],
        "clocks": clk,
        "tally_check": {"records": float(tally_host[capi.TALLY_COUNT:capi.TALLY_COUNT + 7].sum()),
                        "expected": tot["n_particles"] / K if world == 1 else None},
        "single_stream": single,
        "cpu_baseline": cpu,
    }
    if dark:
        line["dark"] = {"dark_vectors_per_shower": tot["n_dark"] / (K * n), "dark_trials_per_sample": tot["n_dark_trials"] / max(tot["n_dark_samples"], 1),
                        "weight_sum_per_shower": float(dtally_host[capi.TALLY_WSUM + 5] / n_job), "mV_used": sh._mV}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS), help="BASELINE.json configuration (1-based); 2 is the headline")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"], help="weak: primaries per GPU fixed; strong: the configuration's primaries split over the GPUs")
    ap.add_argument("--primaries", type=int, default=0, help="primaries per GPU per step (weak) / in total (strong); 0 = the configuration's own number")
    ap.add_argument("--cpu-showers", type=int, default=0, help="size of the CPU-baseline sample (0 = per configuration: about 15-30 s of CPU work on all host cores)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--parts", type=int, default=0, help="sub-batches of a step stepped concurrently on their own engine handle / stream / host thread "
                    "(1 = one batch, one stream; 0 = the configuration's measured best: 1 for the wide-wave configurations 2-4, 2 for the narrow-wave ones 1 and 5)")
    ap.add_argument("--no-history", dest="history", action="store_false", help="skip the e2e_history step (full particle history copied to the host)")
    ap.add_argument("--no-fudge-line", dest="fudge_line", action="store_false", help="skip the maxF_fudge_global=4 line")
    ap.add_argument("--no-retrained-line", dest="retrained_line", action="store_false", help="skip the step on maps retrained in this run")
    args = ap.parse_args()
    if args.impl == "reference":
        reference_arm(args)
    else:
        ours(args)


if __name__ == "__main__":
    main()
