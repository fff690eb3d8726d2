"""Algorithmic work model of the shower step, used by bench.py for the roofline numbers.

Bytes are the HBM traffic the algorithm needs per unit with this record layout (tables and map grids are
L2/shared-memory traffic and are excluded); flops are hand-counted fp64 operations of the reference formulas
(+, -, *, /, sqrt = 1 flop each; cos/sin/exp/log/pow/atan/acos are counted as SPECIAL, not as flops).
"""

# ---- bytes per unit -------------------------------------------------------------------------------------------
RECORD_BYTES = 160                      # pb_stack record: 4 x double4 + key(8) + meta(16) + aux(8)
# k_loop, per charged track: read order(4) p0(32) r0w(32) key(8) meta(16); write pf(32) rf(32) aux(8)
BYTES_LOOP = 4 + 32 + 32 + 8 + 16 + 32 + 32 + 8
# k_finalize, per record: read order(4) meta(16) key(8) + state(64 or 72); write pf(32) rf(32) aux(8) bucket(4)
BYTES_FINALIZE = 4 + 16 + 8 + 72 + 32 + 32 + 8 + 4
BYTES_PROPAGATE = BYTES_FINALIZE          # per record; k_loop adds BYTES_LOOP per charged track
# k_bucket_fill: read bucket(4); write sorted(4)
BYTES_FILL = 8
# k_sample, per accepted sample: read sorted(4) pf.E(8) key(8); write xs(32) ntrials(4)
BYTES_SAMPLE = 4 + 8 + 8 + 32 + 4
# k_emit, per parent: read sorted(4) bucket(4) pf(32) rf(32) weight(8) key(8) meta(16) xs(32); per daughter written:
# p0(32) r0w(32) key(8) meta(16)
BYTES_EMIT_PARENT = 4 + 4 + 32 + 32 + 8 + 8 + 16 + 32
BYTES_EMIT_DAUGHTER = 32 + 32 + 8 + 16


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
# integrand evaluation (reference formula op count) per process; + map transform (5 per dimension) + accept test (3)
FLOPS_INTEGRAND = {"Brem": 100, "MuonBrem": 100, "PairProd": 100, "Comp": 60, "Ann": 35, "Moller": 40, "Bhabha": 60,
                   "MuonE": 35, "DarkBrem": 300, "DarkMuonBrem": 300, "DarkAnn": 80, "DarkComp": 70}
DIM = {"Brem": 4, "MuonBrem": 4, "PairProd": 4, "DarkBrem": 3, "DarkMuonBrem": 3}
SPECIAL_INTEGRAND = {"Brem": 1, "MuonBrem": 1, "PairProd": 1, "DarkBrem": 2, "DarkMuonBrem": 2, "DarkAnn": 5}
FLOPS_SUBSTEP = 120       # 2-3 table interpolations, exp, energy loss, advance, Lynch-Dahl width, two rotations
SPECIAL_SUBSTEP = 10      # exp, log, 2 atan, 4 sincos pairs, Box-Muller log + sincos
FLOPS_KINEMATICS = 150    # sample -> two four-vectors -> lab frame
SPECIAL_KINEMATICS = 8


def flops_per_trial(process):
    return FLOPS_INTEGRAND[process] + 5 * DIM.get(process, 1) + 3


def run_flops(trials_by_process, n_substeps, n_samples):
    f = sum(flops_per_trial(p) * n for p, n in trials_by_process.items())
    return f + FLOPS_SUBSTEP * n_substeps + FLOPS_KINEMATICS * n_samples


def sample_kernel_flops(trials_by_process):
    return sum(flops_per_trial(p) * n for p, n in trials_by_process.items())
