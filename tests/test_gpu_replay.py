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
