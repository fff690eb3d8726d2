#!/usr/bin/env python
"""Headline benchmark: showers/s (and particle-steps/s) of the shower-stepping hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--primaries P]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path over one batch of synthetic primaries: BASELINE.json configs[1], 10 GeV photons
into lead, E_min = 10 MeV, 1e5 primaries per GPU (weak scaling: every rank steps its own 1e5 showers, shower ids
offset by rank; the only collective is one NCCL all-reduce of the 8 KB tally buffer per step).

One JSON line on rank 0:  value = whole-job showers/s with primaries resident in HBM; e2e = the same through the public
host API (pinned host primaries copied in, tallies read back); roofline = the dominant kernel (k_sample) against the
measured HBM peak as the contract asks, and "fp64" = the same kernel against a measured FP64 FMA peak (the path is
FP64-pipe bound); cpu_baseline = the CPU oracle (a port of the reference's generate_shower) on the host
cores over a bounded sample.  ``--impl reference`` times that CPU path alone.
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

MATERIAL, PID, E0, EMIN, SEED = "lead", 22, 10.0, 0.010, 20261017
WORKLOAD = "SM shower: 10 GeV photon into lead, E_min=0.010 GeV, 1e5 primaries (BASELINE.json configs[1])"


# ------------------------------------------------------------------------------------------------ CPU arm
_ORC = None


def _cpu_init():
    global _ORC
    from oracle.shower import OracleShower
    _ORC = OracleShower(None, MATERIAL, EMIN, seed=SEED, rng="counter")


def _cpu_one(i):
    from oracle.shower import OParticle
    sh = _ORC.generate_shower(OParticle([E0, 0.0, 0.0, E0], [0, 0, 0], PID=PID, ID=1, mass=0.0), shower_id=i)
    steps = sum(1 for q in sh if q.ended)
    return len(sh), steps


def cpu_run(n_showers, cores, first_id=0, pool=None):
    """Oracle showers over a process pool -> (seconds, particles, steps)."""
    import multiprocessing as mp
    own = pool is None
    if own:
        pool = mp.get_context("fork").Pool(cores, initializer=_cpu_init)
        pool.map(_cpu_one, range(cores))            # construct the tables in every worker before timing
    t0 = time.perf_counter()
    res = pool.map(_cpu_one, range(first_id, first_id + n_showers), chunksize=1)
    dt = time.perf_counter() - t0
    if own:
        pool.close()
    return dt, sum(r[0] for r in res), sum(r[1] for r in res)


def reference_arm(args):
    """--impl reference: the reference's CPU algorithm (oracle port; the reference itself cannot be installed here,
    see DESIGN.md) on all host cores, same metric and config; each step is a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    pool = mp.get_context("fork").Pool(cores, initializer=_cpu_init)
    pool.map(_cpu_one, range(cores))
    n = args.cpu_showers or max(8 * cores, 64)
    for w in range(args.warmup):
        cpu_run(n, cores, first_id=10_000 + w * n, pool=pool)
    tot_t, tot_p, tot_s = 0.0, 0, 0
    for k in range(args.steps):
        dt, npart, nsteps = cpu_run(n, cores, first_id=k * n, pool=pool)
        tot_t += dt; tot_p += npart; tot_s += nsteps
    pool.close()
    v = args.steps * n / tot_t
    line = {"impl": "reference", "metric": "showers/sec", "value": v, "unit": "showers/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "material": MATERIAL, "pid": PID, "E0_GeV": E0, "E_min_GeV": EMIN,
                       "showers_per_step": n, "note": "CPU oracle port of PETITE generate_shower (vectorised sweeps), "
                                                      "multiprocessing over all host cores"},
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

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        # before CUDA is initialised in this process (fork-safe)
        cores = os.cpu_count() or 1
        n_cpu = args.cpu_showers or max(8 * cores, 64)
        dt, npart, nsteps = cpu_run(n_cpu, cores)
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
    sh = Shower(DATA, MATERIAL, EMIN, seed=SEED, device=local_rank)
    n = args.primaries

    def host_primaries(k):
        pin = lambda shape, dt: torch.empty(shape, dtype=dt, pin_memory=True)
        p = pin((k, 4), torch.float64); p[:] = torch.tensor([E0, 0.0, 0.0, E0], dtype=torch.float64)
        r = pin((k, 3), torch.float64); r.zero_()
        w = pin((k,), torch.float64); w.fill_(1.0)
        m = pin((k,), torch.float64); m.zero_()
        pid = pin((k,), torch.int32); pid.fill_(PID)
        fl = pin((k,), torch.int32); fl.zero_()
        return [p, r, w, m, pid, fl]

    host = host_primaries(n)
    host_np = [t.numpy() for t in host]
    devp = [t.to(dev) for t in host]
    h2d_bytes = sum(t.numel() * t.element_size() for t in host)

    # calibration: records per shower and widest wave -> stack capacity in HBM
    ncal = min(n, 2000)
    cal = sh.run_arrays(*[a[:ncal] for a in host_np], first_shower_id=10 ** 9)
    per = cal.n / ncal
    capacity = int(n * per * 1.06 + 2.3 * cal.counters["max_wave"] / ncal * n) + (1 << 16)
    sh._ensure_stack(capacity)
    tally = torch.zeros(capi.TALLY_SIZE, dtype=torch.float64, device=dev)
    base_id, _ = shard(world * n, rank, world)          # weak scaling: rank r steps global showers [r*n, (r+1)*n)

    P = max(1, args.parts)

    def step(first_id, arrays, parts=1):
        tally.zero_()
        if parts > 1:                     # concurrent sub-batches: one engine handle, stream and host thread per part,
            bs = sh.run_arrays_split(*arrays, parts=parts, capacity=capacity, first_shower_id=first_id, tally=tally)   # tallied per part
        else:
            bs = [sh.run_arrays(*arrays, capacity=capacity, first_shower_id=first_id)]
            sh.tally_batches(bs, tally)
        if world > 1:
            dist.all_reduce(tally)        # the only collective: 8 KB of tallies over NVLink
        return bs

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # warm-up (single stream; in split mode also the peer engines: tables, stacks, scratch growth); the last single-stream
    # warm-up step times EVERY kernel (profiling level 2, ~6 % overhead) for the per-kernel table
    full_ms, full_launch = {}, {}
    for w in range(args.warmup):
        if w == args.warmup - 1:
            sh.set_profiling(2)
        step(base_id, devp)
    if args.warmup:
        pr = sh.get_profile()
        full_ms, full_launch = dict(pr["ms"]), dict(pr["launches"])
    sh.set_profiling(0)
    if P > 1:
        for w in range(max(args.warmup, 1)):
            step(base_id, devp, P)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    COUNT_KEYS = ("n_particles", "n_steps", "n_substeps", "n_samples", "n_trials", "n_launches", "n_waves", "n_charged")

    def timed_region(parts, level):
        """K steps bracketed by barrier + synchronize; -> (ms max over ranks, summed counters, per-kernel ms / launches, trials)."""
        sh.set_profiling(level)
        pm = {k: 0.0 for k in capi.KERNEL_NAMES}
        pl = {k: 0 for k in capi.KERNEL_NAMES}
        tr = {}
        tt = {k: 0 for k in COUNT_KEYS}
        barrier()
        e0.record()
        for k in range(args.steps):
            bs = step(base_id, devp, parts)
            if level:
                pr = sh.get_profile()
                for name in pm:
                    pm[name] += pr["ms"][name]; pl[name] += pr["launches"][name]
                for p_, v in pr["trials"].items():
                    tr[p_] = tr.get(p_, 0) + v
            for b in bs:
                for key in tt:
                    tt[key] += b.counters[key]
            tt["n_launches"] += len(bs) + (1 if world > 1 else 0)     # k_tally per part (+ NCCL kernel)
        e1.record()
        barrier()
        t_ms = e0.elapsed_time(e1)
        sh.set_profiling(0)
        if world > 1:
            t = torch.tensor([t_ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            t_ms = float(t.item())
        return t_ms, tt, pm, pl, tr

    # ---- timed region: K steps, device-resident primaries.  With --parts P > 1 (default 2) a step runs as P concurrent
    # sub-batches and carries no per-kernel events; the single-stream pass after it (same K steps, same inputs) times the two
    # dominant kernels with CUDA events (level 1) while each launch owns the GPU: that is what the roofline is defined on.
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    ms, tot, prof_ms, prof_launch, trials = timed_region(P, 0 if P > 1 else 1)
    clk = clocks.stop() if rank == 0 else None
    single = None
    if P > 1:
        ms1, tot1, prof_ms, prof_launch, trials = timed_region(1, 1)
        single = {"value": world * n * args.steps / (ms1 * 1e-3), "unit": "showers/s", "ms_per_step": ms1 / args.steps,
                  "note": "the same K steps as ONE batch on one stream: the region the per-kernel times, roofline and fp64 figures come from"}
        launches_main = tot["n_launches"]
        tot = dict(tot1, n_launches=launches_main)          # identical showers: only the launch count differs
    tally_host = tally.cpu().numpy()

    # ---- end-to-end through the public host API: pinned host primaries in, tallies out, every step
    e2e_steps = max(1, min(args.steps, 3))
    step(base_id, host_np, P)            # untimed: host-staging buffers of every engine handle at full size
    barrier()
    e0.record()
    for k in range(e2e_steps):
        step(base_id, host_np, P)
        _ = tally.cpu()
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms_e2e], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_e2e = float(t.item())

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    hbm_peak, peak_src = measured_peaks()
    fp64_peak = sh.measure_fp64_peak()
    K = args.steps
    tot["n_daughters"] = tot["n_particles"] - K * n
    for k in prof_ms:                                           # kernels not timed live: scale the warm-up measurement
        if prof_ms[k] == 0.0 and full_ms.get(k):
            prof_ms[k] = full_ms[k] * K
            prof_launch[k] = prof_launch[k] or full_launch[k] * K
    step_ms_total = sum(prof_ms.values())
    dom = max(("k_loop", "k_sample"), key=prof_ms.get)          # dominant kernel of the step (both timed live)
    dom_ms = prof_ms[dom]
    dom_bytes = rl.kernel_bytes(dom, tot)
    ach = dom_bytes / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
    dom_flops = rl.kernel_flops(dom, tot, trials)
    all_bytes = rl.step_bytes(tot["n_particles"], tot["n_charged"], tot["n_samples"], tot["n_daughters"])
    per_kernel = {k: {"ms_per_step": prof_ms[k] / K, "launches_per_step": prof_launch[k] / K,
                      "alg_GBps": (rl.kernel_bytes(k, tot) / (prof_ms[k] * 1e-3) / 1e9) if prof_ms[k] > 0 else 0.0,
                      "model_fp64_tflops": (rl.kernel_flops(k, tot, trials) / (prof_ms[k] * 1e-3) / 1e12) if prof_ms[k] > 0 else 0.0}
                  for k in prof_ms}
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r01_traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get(dom)
        except Exception:
            traffic = None
    line = {
        "metric": "showers/sec", "value": world * n * K / (ms * 1e-3), "unit": "showers/s", "n_gpus": world, "steps": K,
        "warmup": args.warmup, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "material": MATERIAL, "pid": PID, "E0_GeV": E0, "E_min_GeV": EMIN,
                   "primaries_per_gpu": n, "parallelism": f"{world} x independent shower shards, tallies all-reduced" + (f"; each shard stepped as {P} concurrent sub-batches (streams)" if P > 1 else ""),
                   "seed": SEED, "maxF": "regenerated (oracle.findmax, B=300)",
                   "l2": f"working set {capacity * rl.RECORD_BYTES / 1e9:.1f} GB of stack per GPU >> 126 MB L2 (no flush needed)",
                   "stack_capacity_records": capacity},
        "particle_steps_per_sec": world * tot["n_steps"] / (ms * 1e-3),
        "trials_per_sec": world * tot["n_trials"] / (ms * 1e-3),
        "per_shower": {"records": tot["n_particles"] / (K * n), "steps": tot["n_steps"] / (K * n),
                       "substeps": tot["n_substeps"] / (K * n), "trials": tot["n_trials"] / (K * n),
                       "waves": tot["n_waves"] / K},
        "roofline": {"kernel": dom, "bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s",
                     "frac": ach / hbm_peak, "traffic": traffic["dram_bytes_per_launch"] if traffic else None,
                     "traffic_capture": traffic, "peak_source": peak_src,
                     "launches": prof_launch[dom], "avg_launch_ms": dom_ms / max(prof_launch[dom], 1),
                     "algorithmic_bytes_per_launch": dom_bytes / max(prof_launch[dom], 1),
                     "share_of_step": dom_ms / step_ms_total if step_ms_total else None,
                     "timed_region": ("single-stream pass (single_stream): the main region runs %d sub-batches concurrently and its launches overlap" % P) if P > 1 else "main",
                     "note": "the step is FP64-pipe / divergence bound, not HBM bound (SURVEY.md 8d): see fp64"},
        "fp64": {"kernel": dom, "achieved_tflops": dom_flops / (dom_ms * 1e-3) / 1e12 if dom_ms else 0.0,
                 "peak_tflops": fp64_peak, "frac": (dom_flops / (dom_ms * 1e-3) / 1e12 / fp64_peak) if dom_ms and fp64_peak else None,
                 "peak_source": "pb_measure_fp64_peak (DFMA chains, this run)",
                 "model": "hand-counted reference flops per unit (petite_b200/roofline.py) x measured counters; specials not counted"},
        "whole_step_hbm": {"algorithmic_GBps": all_bytes / (ms * 1e-3) / 1e9, "frac_of_peak": all_bytes / (ms * 1e-3) / 1e9 / hbm_peak},
        "kernels": per_kernel,
        "trials_by_process": trials,
        "e2e": {"value": world * n * e2e_steps / (ms_e2e * 1e-3), "unit": "showers/s", "h2d_bytes_per_step": h2d_bytes,
                "d2h_bytes_per_step": int(tally_host.nbytes + 8 * 16), "steps": e2e_steps,
                "api": ("Shower.run_arrays_split(host arrays, parts=%d)" % P if P > 1 else "Shower.run_arrays(host arrays)") + " + pb_tally + tally.cpu()"},
        "gpu_launches": tot["n_launches"],
        "clocks": clk,
        "tally_check": {"records": float(tally_host[capi.TALLY_COUNT:capi.TALLY_COUNT + 7].sum()),
                        "expected": world * tot["n_particles"] / K},
        "single_stream": single,
        "cpu_baseline": cpu,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--primaries", type=int, default=100_000, help="primaries per GPU per step")
    ap.add_argument("--cpu-showers", type=int, default=0, help="size of the CPU-baseline sample (0 = 8 x cores: about 15-20 s of CPU work on all host cores)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--parts", type=int, default=2, help="sub-batches of a step stepped concurrently on their own engine handle / stream / host thread (1 = one batch, one stream)")
    args = ap.parse_args()
    if args.impl == "reference":
        reference_arm(args)
    else:
        ours(args)


if __name__ == "__main__":
    main()
