"""Algorithmic work model of the shower step, used by bench.py for the roofline numbers.

Bytes are the HBM traffic the algorithm needs per unit with this record layout (tables and map grids are
L2/shared-memory traffic and are excluded); flops are hand-counted fp64 operations of the reference formulas
(+, -, *, /, sqrt = 1 flop each; cos/sin/exp/log/pow/atan/acos are counted as SPECIAL, not as flops).
"""

# ---- bytes per unit -------------------------------------------------------------------------------------------
RECORD_BYTES = 168                      # pb_stack record: 4 x double4 + ids(32: meta, key, weight) + aux(8)
# k_loop, per charged track: read order(4) p0(32) r0w(32) track set-up(16) ids(32); write pf(32) rf(32) aux(8)
BYTES_LOOP = 4 + 32 + 32 + 16 + 32 + 32 + 32 + 8
# k_finalize, per record: read order(4) ids(32) + state(64 or 72); write pf(32) rf(32) aux(8) bucket(4)
BYTES_FINALIZE = 4 + 32 + 72 + 32 + 32 + 8 + 4
BYTES_PROPAGATE = BYTES_FINALIZE          # per record; k_loop adds BYTES_LOOP per charged track
# k_bucket_fill, per record: read bucket(4), write sorted(4); per sampled record also read E(8) key(8), write sE(8) skey(8)
BYTES_FILL = 8 + 32
# k_sample, per accepted sample: read sorted(4) sE(8) skey(8); write xs(32) ntrials(4)
BYTES_SAMPLE = 4 + 8 + 8 + 32 + 4
# k_emit, per parent: read sorted(4) bucket(4) pf(32) rf(32) ids(32) xs(32); per daughter written: p0(32) r0w(32) ids(32)
BYTES_EMIT_PARENT = 4 + 4 + 32 + 32 + 32 + 32
BYTES_EMIT_DAUGHTER = 32 + 32 + 32


def step_bytes(n_records, n_charged, n_samples, n_daughters):
    """Algorithmic HBM bytes of a whole run."""
    return (n_records * (BYTES_FINALIZE + BYTES_FILL) + n_charged * BYTES_LOOP
            + n_samples * (BYTES_SAMPLE + BYTES_EMIT_PARENT) + n_daughters * BYTES_EMIT_DAUGHTER)


def kernel_bytes(kernel, c):
    """Algorithmic HBM bytes of one kernel over a run with counters ``c`` (summed pb_counters) and n_daughters."""
    return {"k_loop": c["n_charged"] * BYTES_LOOP, "k_finalize": c["n_particles"] * BYTES_FINALIZE,
            "k_bucket_fill": c["n_particles"] * BYTES_FILL, "k_sample": c["n_samples"] * BYTES_SAMPLE,
            "k_emit": c["n_samples"] * BYTES_EMIT_PARENT + c["n_daughters"] * BYTES_EMIT_DAUGHTER}.get(kernel, 0)


def kernel_flops(kernel, c, trials_by_process):
    return {"k_loop": c["n_substeps"] * FLOPS_SUBSTEP, "k_finalize": c["n_charged"] * FLOPS_SUBSTEP + c["n_particles"] * 40,
            "k_sample": sample_kernel_flops(trials_by_process), "k_emit": c["n_samples"] * FLOPS_KINEMATICS}.get(kernel, 0)


# ---- fp64 flops per unit --------------------------------------------------------------------------------------
# COUNTED, not estimated: `python -m oracle.count_ops` runs the oracle's restatement of the reference formulas on operation-
# counting number types (ndarray subclass for the vectorised integrands, float subclass + math proxy for the scalar pieces)
# and prints these constants.  +, -, *, /, sqrt = 1 flop each; cos/sin/exp/log/pow/atan/acos = 1 SPECIAL each, counted
# separately and NOT converted into flops.  Per accept/reject trial: integrand at one point + map transform + accept test.
FLOPS_TRIAL = {"Brem": 108, "Ann": 18, "PairProd": 108, "Comp": 31, "Moller": 35, "Bhabha": 62, "MuonE": 24, "MuonBrem": 108,
               "DarkBrem": 226, "DarkAnn": 26, "DarkComp": 30, "DarkMuonBrem": 226}
SPECIAL_TRIAL = {"Brem": 1, "Ann": 0, "PairProd": 1, "Comp": 0, "Moller": 0, "Bhabha": 0, "MuonE": 0, "MuonBrem": 1,
                 "DarkBrem": 3, "DarkAnn": 3, "DarkComp": 0, "DarkMuonBrem": 3}
FLOPS_SUBSTEP = 148       # one dE/dx + multiple-scattering sub-step (shower.py:559-581): 3 table interpolations, energy loss, advance,
SPECIAL_SUBSTEP = 16      # Lynch-Dahl width, rotation there and back; specials: exp, 2 log, 2 atan, pow, 5 cos + 5 sin
FLOPS_KINEMATICS = 95     # sampled point -> two four-vectors -> rotation matrix -> lab frame (brem)
SPECIAL_KINEMATICS = 12


def flops_per_trial(process):
    return FLOPS_TRIAL[process]


def run_flops(trials_by_process, n_substeps, n_samples):
    f = sum(flops_per_trial(p) * n for p, n in trials_by_process.items())
    return f + FLOPS_SUBSTEP * n_substeps + FLOPS_KINEMATICS * n_samples


def sample_kernel_flops(trials_by_process):
    return sum(flops_per_trial(p) * n for p, n in trials_by_process.items())


def sample_kernel_specials(trials_by_process):
    return sum(SPECIAL_TRIAL[p] * n for p, n in trials_by_process.items())
