"""Whole showers on the GPU: replay parity with the CPU oracle, invariances, conservation, ensemble statistics."""
import numpy as np
import pytest

from tests.conftest import DATA as DATA_DIR
from tests.gpu_util import shower, primaries
from tests.parity import compare_with_oracle, oracle_showers

pytestmark = pytest.mark.gpu

CASES = [  # material, pid, E, Emin, n
    ("graphite", 11, 10.0, 0.010, 3),       # BASELINE config 1 (README example)
    ("lead", 22, 10.0, 0.010, 2),           # BASELINE config 2
    ("lead", -11, 3.0, 0.010, 4),
    ("graphite", 22, 1.0, 0.010, 16),
    ("graphite", 13, 20.0, 0.030, 2),       # muons: MuonBrem/MuonE + Q-12
    ("lead", -13, 5.0, 0.030, 2),
]


@pytest.mark.parametrize("material,pid,E,Emin,n", CASES)
def test_replay_parity_with_oracle(material, pid, E, Emin, n):
    """Counter-mode replay: same Philox draws on both sides -> identical multiplicities, PIDs, process choices,
    accept/reject trial counts and sub-step counts; four-vectors agree up to amplified rounding."""
    sh = shower(material, Emin, seed=11)
    prims = primaries(pid, E, n)
    batch = sh.generate_showers(prims, first_shower_id=1000)
    rep = compare_with_oracle(batch, prims, material, Emin, seed=11)
    assert rep["structure_mismatch"] == 0, rep
    assert rep["particles"] == batch.n
    assert rep["max_rel_p0"] < 1e-6 and rep["max_rel_pf"] < 1e-6 and rep["max_abs_rf"] < 1e-6, rep


@pytest.mark.parametrize("material,pid,E,Emin,n", [("graphite", 211, 20.0, 0.030, 3), ("lead", -211, 5.0, 0.030, 3),
                                                   ("graphite", 321, 30.0, 0.030, 2), ("lead", -321, 8.0, 0.010, 2)])
def test_long_lived_decay_in_flight_parity_with_oracle(material, pid, E, Emin, n):
    """Row f-5: pi+- / K+- primaries with stability 'long-lived' decay in flight to mu nu (particle.py:363-389, 410-422): decay
    point drawn by accept/reject, the two daughters weighted by BR x P(decay before interaction) at independently drawn distances,
    then the muon showers.  Counter-mode replay against the oracle, which reproduces the unmodified reference's own long-lived
    showers in stream mode (tests/test_oracle_golden.py::test_stream_mode_long_lived_shower_equals_reference)."""
    sh = shower(material, Emin, seed=21)
    prims = primaries(pid, E, n, stability="long-lived")
    batch = sh.generate_showers(prims, first_shower_id=40)
    rep = compare_with_oracle(batch, prims, material, Emin, seed=21)
    assert rep["structure_mismatch"] == 0, rep
    assert rep["particles"] == batch.n and batch.n > 20 * n
    assert rep["max_rel_p0"] < 1e-6 and rep["max_rel_pf"] < 1e-6 and rep["max_abs_rf"] < 1e-6 and rep["max_rel_weight"] < 1e-12, rep
    out = batch.to_particles(prims)
    for plist in out:
        parent, mu, nu = plist[0], plist[1], plist[2]
        assert mu.get_ids()["PID"] == (-13 if pid > 0 else 13) and nu.get_ids()["PID"] == (14 if pid > 0 else -14)
        assert mu.get_ids()["generation_process"] == "SMDecay" and mu.get_ids()["parent_ID"] == -1                     # Q-10
        br = 0.9998 if abs(pid) == 211 else 0.6356
        assert 0 < mu.get_ids()["weight"] < br and 0 < nu.get_ids()["weight"] < br and mu.get_ids()["weight"] != nu.get_ids()["weight"]
        assert parent.get_rf()[2] > 0 and np.allclose(mu.get_r0(), parent.get_rf(), rtol=0, atol=0)                    # daughters start at the decay point
        assert abs(mu.get_p0()[0] + nu.get_p0()[0] - E) < 1e-9 * E
    with pytest.raises(ValueError):
        sh.generate_showers(primaries(111, 5.0, 1, stability="long-lived"))


def test_pi0_primaries_and_q7_mass():
    sh = shower("graphite", 0.010, seed=3)
    from petite_b200 import Particle
    from petite_b200.constants import m_pi0, m_electron
    prims = [Particle([8.0, 0.3, -0.2, np.sqrt(64 - m_pi0 ** 2 - 0.13)], [0, 0, 0.1], {"PID": 111, "ID": 1, "mass": m_pi0, "stability": "short-lived"}),
             Particle([2.0, 0, 0, np.sqrt(4 - m_electron ** 2)], [0, 0, 0], {"PID": 11, "ID": 0})]     # no mass -> 0.000511 (Q-7), ID 0 (Q-8)
    batch = sh.generate_showers(prims, first_shower_id=0)
    rep = compare_with_oracle(batch, prims, "graphite", 0.010, seed=3)
    assert rep["structure_mismatch"] == 0, rep
    out = batch.to_particles(prims)
    assert out[0][1].get_ids()["generation_process"] == "SMDecay" and out[0][1].get_ids()["parent_ID"] == -1    # Q-10
    assert abs(out[0][1].get_ids()["weight"] - 0.98823) < 1e-15
    assert out[1][0].get_ids()["mass"] == 0.000511 and out[1][1].get_ids()["ID"] in (0, 1)


def test_generate_shower_single_primary_api():
    sh = shower("graphite", 0.010, seed=5)
    p0 = primaries(11, 5.0, 1)[0]
    sh._next_shower_id = 77
    out = sh.generate_shower(p0)
    ref = oracle_showers([p0], "graphite", 0.010, 5, first_shower_id=77)[0]
    assert len(out) == len(ref)
    assert [p.get_ids()["PID"] for p in out] == [q.PID for q in ref]
    assert [p.get_ids()["ID"] for p in out] == [q.ID for q in ref]
    assert [p.get_ids()["generation_process"] for p in out] == [q.process for q in ref]
    assert all(p.get_ended() for p in out)
    low = primaries(11, 0.005, 1)[0]
    assert len(sh.generate_shower(low)) == 1


def test_batch_split_and_determinism():
    """A shower depends only on (seed, shower id): one call of 64 == two calls of 32, and reruns are identical."""
    sh = shower("lead", 0.010, seed=9)
    prims = primaries(22, 2.0, 64)

    def signature(batch):
        h = batch.to_host()
        order, offs = batch.reference_order()
        return [np.concatenate([h["p0"][order[offs[i]:offs[i + 1]]].ravel(), h["rf"][order[offs[i]:offs[i + 1]]].ravel()]) for i in range(batch.n_primaries)]
    a = signature(sh.generate_showers(prims, first_shower_id=500))
    b = signature(sh.generate_showers(prims, first_shower_id=500))
    c1 = signature(sh.generate_showers(prims[:32], first_shower_id=500))
    c2 = signature(sh.generate_showers(prims[32:], first_shower_id=532))
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    for x, y in zip(a, c1 + c2):
        assert np.array_equal(x, y)


def test_concurrent_sub_batches_match_single_batch():
    """run_arrays_split: two engine handles, two streams, two host threads - same showers as one call, and the same tallies."""
    sh = shower("lead", 0.010, seed=9)
    prims = primaries(22, 2.0, 65)
    arrays = sh._pack_primaries(prims)

    def signature(batch):
        h = batch.to_host()
        order, offs = batch.reference_order()
        return [np.concatenate([h["p0"][order[offs[i]:offs[i + 1]]].ravel(), h["rf"][order[offs[i]:offs[i + 1]]].ravel()]) for i in range(batch.n_primaries)]
    one = sh.run_arrays(*arrays, first_shower_id=500)
    a = signature(one)
    t_one = sh.tally(one).cpu().numpy()
    parts = sh.run_arrays_split(*arrays, parts=2, first_shower_id=500)
    assert [b.n_primaries for b in parts] == [33, 32] and parts[0]._owner is not parts[1]._owner
    b = signature(parts[0]) + signature(parts[1])
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    t_two = sh.tally_batches(parts).cpu().numpy()
    assert np.allclose(t_one, t_two, rtol=1e-12, atol=0)
    import torch
    t_in = torch.zeros_like(sh.tally(parts[0]))
    sh.run_arrays_split(*arrays, parts=2, first_shower_id=500, tally=t_in)       # tallied inside, per part and stream
    assert np.allclose(t_one, t_in.cpu().numpy(), rtol=1e-12, atol=0)
    assert sum(p.n for p in parts) == one.n
    # device-resident primaries, three parts
    import torch
    dev = [torch.from_numpy(np.ascontiguousarray(x)).cuda() for x in arrays]
    parts3 = sh.run_arrays_split(*dev, parts=3, capacity=one.n * 2 + 4096, first_shower_id=500)
    assert sum(p.n for p in parts3) == one.n


def test_energy_accounting_full_size_property():
    """Size-independent property at a BASELINE-like size: daughters never carry more energy than the parent had at
    the vertex (up to the rest mass of the struck atomic electron), and weights are inherited."""
    sh = shower("lead", 0.010, seed=2)
    prims = primaries(22, 10.0, 2000)
    batch = sh.generate_showers(prims)
    h = batch.to_host()
    par = h["parent"]
    d = par >= 0
    Ed = np.zeros(batch.n)
    np.add.at(Ed, par[d], h["p0"][d][:, 0])
    from petite_b200.constants import m_electron
    assert np.all(Ed <= h["pf"][:, 0] + m_electron + 1e-9)
    assert np.all(h["p0"][d][:, 0] > 0.010)
    assert np.array_equal(h["weight"][d], h["weight"][par[d]])
    assert batch.counters["n_steps"] <= batch.n and batch.counters["n_no_sample"] == 0
    assert np.all(np.isfinite(h["pf"])) and np.all(np.isfinite(h["rf"]))


def test_ensemble_statistics_vs_oracle():
    """KS tests (p > 0.01) of GPU showers against an independent oracle sample (different shower ids)."""
    from scipy.stats import ks_2samp
    sh = shower("graphite", 0.010, seed=21)
    n_gpu, n_orc = 4000, 150
    prims = primaries(11, 1.0, n_gpu)
    batch = sh.generate_showers(prims, first_shower_id=10_000)
    h = batch.to_host()
    mult_gpu = np.bincount(h["shower"], minlength=n_gpu)
    ref = oracle_showers(prims[:n_orc], "graphite", 0.010, 21, first_shower_id=0)
    mult_orc = np.array([len(r) for r in ref])
    assert ks_2samp(mult_gpu, mult_orc).pvalue > 0.01
    ph_gpu = h["p0"][(h["pid"] == 22) & (h["parent"] >= 0)][:, 0]
    ph_orc = np.array([q.p0[0] for r in ref for q in r[1:] if q.PID == 22])
    assert ks_2samp(ph_gpu, ph_orc).pvalue > 0.01
    th = lambda p: np.arccos(np.clip(p[:, 3] / np.linalg.norm(p[:, 1:], axis=1), -1, 1))
    el_gpu = th(h["p0"][(np.abs(h["pid"]) == 11) & (h["parent"] >= 0)])
    el_orc = th(np.array([q.p0 for r in ref for q in r[1:] if abs(q.PID) == 11]))
    assert ks_2samp(el_gpu, el_orc).pvalue > 0.01


def test_edge_cases_empty_ragged_and_invalid():
    """Boundary behaviour: mixed species and energies in one batch (incl. below-threshold primaries that must come back
    untouched), neutrinos, and inputs the reference would hang or crash on."""
    from petite_b200 import Particle
    from petite_b200.constants import m_electron, m_muon
    from petite_b200 import _capi as capi
    sh = shower("graphite", 0.010, seed=8)
    mk = lambda pid, E, m, **kw: Particle([E, 0, 0, np.sqrt(max(E * E - m * m, 0.0))], [0, 0, 0], dict({"PID": pid, "ID": 1, "mass": m}, **kw))
    prims = [mk(11, 0.004, m_electron), mk(22, 0.0012, 0.0), mk(-11, 0.5, m_electron), mk(14, 3.0, 0.0), mk(13, 0.1, m_muon),
             mk(22, 3.0, 0.0), mk(11, 0.0101, m_electron)]
    batch = sh.generate_showers(prims, first_shower_id=40)
    rep = compare_with_oracle(batch, prims, "graphite", 0.010, seed=8)
    assert rep["structure_mismatch"] == 0, rep
    out = batch.to_particles(prims)
    assert len(out[0]) == 1 and np.array_equal(out[0][0].get_pf(), out[0][0].get_p0())      # below min_energy: untouched
    assert len(out[3]) == 1                                                                   # neutrino just ends
    assert len(out[4]) == 1                                                                   # muon below 0.2 GeV threshold
    assert len(out[2]) > 1 and len(out[5]) > 1
    with pytest.raises(ValueError):                                                           # Q-20: would loop forever
        sh.generate_showers([mk(2212, 5.0, 0.938)])
    with pytest.raises(ValueError):
        sh.generate_showers([mk(221, 5.0, 0.547862, stability="short-lived")])               # eta: a three-body channel may be drawn (particle.py:403-404)
    with pytest.raises(capi.EngineError) as ei:                                               # too small a stack fails loudly
        sh.generate_showers(primaries(11, 5.0, 64), capacity=200)
    assert ei.value.code == capi.PB_ERR_CAPACITY
    assert sh.generate_showers(primaries(11, 1.0, 4)).n > 4                                   # engine usable afterwards


def test_no_sample_found_is_reported():
    """max_n_integrators exhausted -> the reference raises "No Sample Found" (shower.py:460-461)."""
    from petite_b200.shower import Shower
    from petite_b200 import _capi as capi
    s = Shower(DATA_DIR, "lead", 0.010, max_n_integrators=1, maxF_fudge_global=1e6, seed=1)
    with pytest.raises(capi.EngineError) as ei:
        s.generate_showers(primaries(22, 5.0, 32))
    assert ei.value.code == capi.PB_ERR_NO_SAMPLE and "No Sample Found" in str(ei.value)


def test_tally_matches_host_histograms():
    from petite_b200 import _capi as capi
    sh = shower("lead", 0.010, seed=12)
    b = sh.generate_showers(primaries(22, 4.0, 300))
    t = sh.tally(b).cpu().numpy()
    h = b.to_host()
    sp = {11: 0, -11: 1, 22: 2}
    for pid, k in sp.items():
        sel = h["pid"] == pid
        assert t[capi.TALLY_COUNT + k] == sel.sum()
        assert abs(t[capi.TALLY_WESUM + k] - np.sum(h["weight"][sel] * h["p0"][sel][:, 0])) < 1e-6 * t[capi.TALLY_WESUM + k]
        eb = np.clip(np.floor((np.log10(h["p0"][sel][:, 0]) + 3) * (64 / 6)).astype(int), 0, 63)
        assert np.array_equal(t[capi.TALLY_EHIST + k * 64: capi.TALLY_EHIST + (k + 1) * 64], np.bincount(eb, minlength=64))
    assert t[capi.TALLY_COUNT:capi.TALLY_COUNT + 7].sum() == b.n


@pytest.mark.parametrize("name,n", [("c1_e_graphite", 1_000), ("c2_gamma_lead", 10_000)])
def test_exact_decisions_at_1e3_1e4_showers(golden, name, n):
    """EXACT equality with the oracle at BASELINE scale: configuration 1 in full (1e3 showers) and the first 1e4 showers of
    configuration 2 (tests/golden/exact_showers.npz, made by the CPU oracle in counter mode, make_exact_fixture.py).  Per shower:
    multiplicity, sum of accept/reject trials, sum of dE/dx sub-steps and the histogram of generation processes.  The hot-loop
    forms (fast_rcp, hot_log, hot_exp_neg_step, mcs_fast, folded integrands, per-species summed n*sigma tables) are <= 2 ulp away
    from the oracle's libm path, so a hard-scatter / process / accept decision CAN flip when a uniform lands within ~1e-15 of a
    threshold; this test measures how often (expected flips at this size: ~1e-7) and demands zero."""
    import torch
    from tests import ensemble_stats as es
    from tests.test_gpu_ensemble import _run
    cfg = es.CONFIGS[name]
    g = golden("exact_showers")
    sh = shower(cfg["material"], cfg["E_min"], seed=cfg["seed"])
    es.REF_CONFIGS.setdefault(name, cfg)
    batch = _run(sh, name, n, first_id=0)
    t, m = batch._t, batch.n
    sid = t["meta"][:m, 3].long()
    mult = torch.bincount(sid, minlength=n).cpu().numpy()
    ntr = torch.bincount(sid, weights=t["aux"][:m, 0].double(), minlength=n).cpu().numpy().astype(np.int64)
    nsub = torch.bincount(sid, weights=t["aux"][:m, 1].double(), minlength=n).cpu().numpy().astype(np.int64)
    proc = (t["meta"][:m, 2] & 0xFF).long()
    hist = torch.bincount(sid * 16 + proc, minlength=16 * n).reshape(n, 16).cpu().numpy()
    bad = (mult != g[f"{name}/mult"]) | (ntr != g[f"{name}/ntrials"]) | (nsub != g[f"{name}/nsub"]) | np.any(hist != g[f"{name}/proc_hist"], axis=1)
    decisions = int(g[f"{name}/ntrials"].sum() + g[f"{name}/nsub"].sum() + 2 * g[f"{name}/mult"].sum())
    print(name, "showers", n, "records", m, "decisions ~", decisions, "showers with any difference:", int(bad.sum()),
          "-> flip rate <", (int(bad.sum()) + 1) / decisions)
    assert not bad.any(), (np.nonzero(bad)[0][:10], mult[bad][:5], g[f"{name}/mult"][bad][:5])


def test_sub_step_cap_does_not_change_the_showers(monkeypatch):
    """k_loop pauses a track when its sub-step index reaches a multiple of PB_LOOP_CAP and carries it into the next wave
    (engine.cu): scheduling only.  The same 400 showers with the cap off, at 4 (almost every track is carried, several times)
    and at the default must be the same records bit for bit - creation and end-point four-vectors, trial and sub-step counts -
    in the reference's order, although the waves they are stepped in differ."""
    from petite_b200.shower import Shower
    prims = primaries(11, 8.0, 400)
    runs = {}
    for cap in ("0", "4", "32"):
        monkeypatch.setenv("PB_LOOP_CAP", cap)
        sh = Shower(DATA_DIR, "lead", 0.010, seed=21)
        b = sh.generate_showers(prims, first_shower_id=300)
        h = b.to_host()
        order, offs = b.reference_order()
        runs[cap] = (b.counters, [h[k][order] for k in ("pid", "p0", "pf", "rf", "ntrials", "nsub", "process")], offs)
        del sh
    c0, a0, o0 = runs["0"]
    for cap in ("4", "32"):
        c, a, o = runs[cap]
        assert np.array_equal(o, o0)
        for x, y in zip(a, a0):
            assert np.array_equal(x, y), cap
        assert all(c[k] == c0[k] for k in ("n_particles", "n_steps", "n_substeps", "n_samples", "n_trials"))
    assert runs["4"][0]["n_waves"] > runs["0"][0]["n_waves"]       # the carried tracks did take extra waves


def test_overwritten_batch_warns():
    """A ShowerBatch is a view of its Shower's reusable stack: reading it after a later run on the same Shower warns (ADVICE, round 1)."""
    import warnings
    sh = shower("graphite", 0.010, seed=4)
    b1 = sh.generate_showers(primaries(11, 1.0, 2), first_shower_id=0)
    h1 = b1.to_host()                                   # copied in time: stays valid and silent
    b2 = sh.generate_showers(primaries(11, 1.0, 3), first_shower_id=10)
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        assert b1.to_host() is h1 and b2.to_host()["pid"].shape[0] == b2.n
    b3 = sh.generate_showers(primaries(11, 1.0, 2), first_shower_id=20)
    b4 = sh.generate_showers(primaries(11, 1.0, 2), first_shower_id=30)
    with pytest.warns(RuntimeWarning):
        b3.to_host()
    with pytest.warns(RuntimeWarning):
        sh.tally(b3)
    assert b4.n > 0
