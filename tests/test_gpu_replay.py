"""Replay mode (BASELINE.json north_star: "a replay mode feeding the reference's uniform stream must give bit-identical process
choices, accept/reject decisions and daughter counts").

tests/golden/replay_tape.npz holds, for five whole showers of the UNMODIFIED reference in stream mode (make_replay_tape.py),
the random numbers each particle-step consumed, in the reference's consumption order, and what the reference made of the step.
``pb_replay`` runs the wave kernels' own device functions (substep, finalize_one, the sampler's folded integrands,
scatter_products) on those tapes.  Demanded, for every one of the ~2 000 particle-steps: the step consumes EXACTLY its tape
segment (same number of sub-steps, same hard-scatter decision, same number of accept/reject trials), picks the same process,
keeps the same daughters with the same PDG ids; four-vectors within the conditioning of the reference's own formulas."""
import numpy as np
import pytest

from tests.gpu_util import shower

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", range(5))
def test_reference_tape_replays_with_identical_decisions(golden, case):
    g = golden("replay_tape")
    pre = f"{case}/"
    pid0, E0, Emin, seed = g[pre + "case"]
    sh = shower(str(g[pre + "material"]), float(Emin))
    part, tape, off, ref = g[pre + "particles"], g[pre + "tape"], g[pre + "tape_off"], g[pre + "ref"]
    n = len(part)
    out = sh.replay(part, tape, off)
    assert n > 100
    status = out["status"].copy()
    # the reference leaves its loop BEFORE the process choice when the last open particle ends below min_energy
    # (shower.py:660-662), so that one step's tape has no choice uniform; the engine draws it (it cannot change anything: the
    # particle is below the sample_scattering threshold) and finds the tape empty exactly there
    last = n - 1
    if status[last] == 1 and out["consumed"][last] == off[last + 1] - off[last] and ref[last, 2] < float(Emin):
        status[last] = 0
    bad = np.nonzero(status != 0)[0]
    assert len(bad) == 0, (case, bad[:10], out["status"][bad[:10]], out["consumed"][bad[:10]], np.diff(off)[bad[:10]], part[bad[:10], 0], ref[bad[:10], :3])
    assert np.array_equal(out["consumed"], np.diff(off))                      # every sub-step, trial and azimuth draw of the reference
    sampled = ref[:, 1] > 0
    assert np.array_equal(out["ntrials"][sampled], ref[sampled, 1].astype(np.int64))     # accept/reject decisions
    assert np.all(out["ntrials"][~sampled] == 0)
    assert np.array_equal(out["process"][sampled], ref[sampled, 0].astype(int))          # process choices
    assert np.array_equal(out["kept"], ref[:, 9].astype(int))                            # daughter counts
    for b, (pk, vk, col) in enumerate((("pid_a", "p_a", 10), ("pid_b", "p_b", 15))):
        has = (ref[:, 9].astype(int) >> b) & 1 == 1
        assert np.array_equal(out[pk][has], ref[has, col].astype(int))
        want = ref[has, col + 1:col + 5]
        err = np.max(np.abs(out[vk][has] - want), axis=1) / np.max(np.abs(want), axis=1)
        # the reference rotates daughters with acos(pz / |p|) of the parent (particle.py:181): a last-bit difference in the
        # parent's direction is amplified by 1 / theta for collimated particles; 1e-8 is what the oracle itself achieves
        assert np.max(err) < 1e-8, (case, b, float(np.max(err)))
    # propagated state: energy to 1e-12, momentum within the multiple-scattering conditioning (gamma^2-amplified in the reference)
    e_err = np.abs(out["pf"][:, 0] - ref[:, 2]) / ref[:, 2]
    assert np.max(e_err) < 1e-12, float(np.max(e_err))
    p_err = np.max(np.abs(out["pf"][:, 1:] - ref[:, 3:6]), axis=1) / np.maximum(np.linalg.norm(ref[:, 3:6], axis=1), 1e-300)
    r_err = np.max(np.abs(out["rf"] - ref[:, 6:9]), axis=1) / np.maximum(1.0, np.max(np.abs(ref[:, 6:9]), axis=1))
    print("case", case, "steps", n, "trials", int(out["ntrials"].sum()), "sub-steps", int(out["nsub"].sum()),
          "max rel dE", float(np.max(e_err)), "max |dp|/|p|", float(np.max(p_err)), "max dr", float(np.max(r_err)))
    assert np.max(p_err) < 1e-8 and np.max(r_err) < 1e-9


def test_replay_detects_a_wrong_tape(golden):
    """Sanity of the check itself: raising the accept uniform of a step's LAST (accepted) trial to ~1 turns the accept into a
    reject, and the step no longer fits its tape (it runs past the end of its segment)."""
    g = golden("replay_tape")
    sh = shower(str(g["1/material"]), 0.010)
    part, tape, off, ref = g["1/particles"], g["1/tape"].copy(), g["1/tape_off"], g["1/ref"]
    k = int(np.nonzero(ref[:, 1] > 0)[0][5])
    tape[off[k + 1] - 2] = 1.0 - 2.0 ** -53          # segment = ..., trials x (y[dim], u), azimuth
    out = sh.replay(part, tape, off)
    assert out["status"][k] == 1 and np.all(np.delete(out["status"], k) == 0)


def test_long_lived_decay_step_replays_a_tape():
    """Tape mode of the decay in flight (row f-5): the tape holds what the reference's generators returned, in its order
    (particle.py:371-372, 412, 420-421, 234-235) - per accept/reject iteration uniform(0, x_max) and uniform(0, 1) for the decay
    point and for each daughter's weight, then cos(theta) = uniform(-1, 1) and the azimuth uniform.  The device step must consume
    exactly that segment (incl. one rejected iteration) and reproduce the oracle's decay fed the same numbers."""
    from oracle.shower import OracleShower, OParticle
    from oracle import consts as OC
    sh = shower("graphite", 0.030)
    pid, E = -211, 12.0
    m = OC.MASS[pid]
    p0 = [E, 0.3, -0.2, np.sqrt(E * E - m * m - 0.13)]
    o = OracleShower(None, "graphite", 0.030, rng="stream")
    prim = OParticle(p0, [0.0, 0.01, 0.02], PID=pid, ID=1, mass=m, stability="long-lived")
    ctau, tot = o._decay_rates(prim, OC.INT_LENGTH[pid], OC.DECAY_LENGTH[pid])
    x_max = 4 / tot
    loops = {1: [(0.9 * x_max, 0.9), (0.31 * x_max, 0.2)], 2: [(0.55 * x_max, 0.05)], 3: [(0.08 * x_max, 0.5)]}    # (x, u): loop 1 rejects once
    assert 0.9 > np.exp(-tot * 0.9 * x_max) and 0.2 < np.exp(-tot * 0.31 * x_max)
    cth, uphi = -0.37, 0.81
    tape = [v for k in (1, 2, 3) for xu in loops[k] for v in xu] + [cth, uphi]

    class Fixed:
        def decay_x(self, loop, i):
            x, u = loops[loop][i]
            return x / x_max, u
        def decay(self, pc=12):
            return 0.5 * (cth + 1.0), uphi
        def child(self, bit):
            return self
    prim.draws = Fixed()
    mu, nu = o.decay(prim)
    part = np.array([[pid, *p0, 0.0, 0.01, 0.02, m, 4.0]])
    out = sh.replay(part, np.array(tape), np.array([0, len(tape)]))
    assert out["status"][0] == 0 and out["consumed"][0] == len(tape) and out["process"][0] == 12
    assert out["pid_a"][0] == 13 and out["pid_b"][0] == -14 and out["kept"][0] == 3
    assert np.allclose(out["rf"][0], prim.rf, rtol=1e-13, atol=0) and prim.rf[2] > 0.02
    assert abs(out["weight_factor"][0] - mu.weight) < 1e-13 * mu.weight and abs(out["weight_factor_b"][0] - nu.weight) < 1e-13 * nu.weight
    assert mu.weight != nu.weight
    assert np.allclose(out["p_a"][0], mu.p0, rtol=1e-11, atol=1e-13) and np.allclose(out["p_b"][0], nu.p0, rtol=1e-11, atol=1e-13)
