"""draw_sample / sample_scattering / find_max through the C ABI."""
import numpy as np
import pytest

from tests.conftest import DATA
from tests.gpu_util import shower, primaries

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("process,E", [("Brem", 1.0), ("PairProd", 10.0), ("Comp", 0.05), ("Moller", 2.0), ("Ann", 0.3),
                                        ("Bhabha", 5.0), ("MuonBrem", 30.0), ("MuonE", 8.0)])
def test_draw_samples_replay_vs_oracle(process, E):
    from oracle.shower import OracleShower
    from oracle.draws import CounterDraws
    from oracle.philox import root_key
    sh = shower("lead", 0.010, seed=17)
    n = 64
    x, ntr = sh.draw_samples(np.full(n, E), process, first_id=300)
    o = OracleShower(None, "lead", 0.010, seed=17)
    for i in range(n):
        xo, no = o.draw_sample(E, process, CounterDraws(root_key(17, 300 + i)))
        assert ntr[i] == no
        assert np.array_equal(x[i], xo)          # same Philox draws, same map arithmetic -> bit-identical sample


def test_draw_sample_and_sample_scattering_api():
    sh = shower("graphite", 0.010, seed=4)
    x = sh.draw_sample(1.0, process="Brem")
    assert x.shape == (4,) and 0 <= x[0] <= 1
    xv = sh.draw_sample(1.0, process="Brem", VB=True)
    assert xv.shape == (5,) and xv[4] >= 1
    fixed = sh.draw_sample(1.0, LU_Key=25, process="Brem")
    assert fixed.shape == (4,)
    with pytest.raises(Exception):
        sh.draw_sample(1.0, process="NoSuchProcess")
    p0 = primaries(22, 10.0, 1)[0]
    d = sh.sample_scattering(p0, "PairProd")
    assert [q.get_ids()["PID"] for q in d] == [-11, 11]
    assert abs(d[0].get_p0()[0] + d[1].get_p0()[0] - 10.0) < 1e-9                    # nucleus takes no energy in this model
    assert [q.get_ids()["ID"] for q in d] == [2, 3] and d[0].get_ids()["generation_process"] == "PairProd"
    low = primaries(11, 0.008, 1)[0]
    assert sh.sample_scattering(low, "Brem") is None
    # tutorial.ipynb:[54]: e+/e- symmetry of pair production
    x, _ = sh.draw_samples(np.full(20000, 10.0), "PairProd")
    asym = (np.sum(x[:, 0] > 0.5) - np.sum(x[:, 0] < 0.5)) / len(x)
    assert abs(asym) < 4 * 0.00707


@pytest.mark.parametrize("material", ["graphite", "lead"])
def test_gpu_find_max_reproduces_shipped_cross_sections(material):
    """find_maxes on the GPU: sigma through the maps must reproduce the shipped sm_xsec rows (the statistical pin of the
    VEGAS-map restatement) and max_F must be compatible with the oracle-built fixture."""
    sh = shower(material, 0.010, seed=0)
    xs = np.load(DATA + "sm_xsec.npz")
    fx = np.load(DATA + "sm_maxF.npz")
    for P in ["Brem", "PairProd", "Comp", "Ann", "Moller", "Bhabha", "MuonE", "MuonBrem"]:
        mf, sg = sh.find_max(P, n_trials=100, seed=5)
        ref = xs[f"{P}/{material}"][:, 1]
        ok = ref > 0
        ratio = sg[ok] / ref[ok]
        assert abs(np.mean(ratio) - 1) < 0.03, (P, np.mean(ratio))
        assert np.all(np.isfinite(mf)) and np.all(mf >= 0)
        want = fx[f"{P}/{material}"]
        pos = want > 0
        r = mf[pos] / want[pos]
        assert np.median(r) > 0.6 and np.median(r) < 1.6, (P, np.median(r))


def test_detector_cut_vs_reference_golden(golden):
    """pb_detector_cut against the reference's detector_cut (shower.py:825-864) on the same particle list."""
    from petite_b200 import Particle
    from petite_b200.analysis import detector_cut
    g = golden("detector")
    sh = shower("graphite", 0.010, seed=4)
    plist = [Particle(list(g["p0"][i]), list(g["r0"][i]), {"PID": 11, "mass": 0.00051099895, "weight": float(g["w"][i])})
             for i in range(len(g["w"]))]
    batch = sh.batch_from_particles(plist)
    z = list(g["z"])
    for tag, kw in (("a", dict(detector_radius=0.5)), ("b", dict(detector_radius=2.0, energy_cut=(1.0, 3.0), detector_inner_radius=0.2))):
        tot = detector_cut(batch, sh, z, method="TotalWeight", **kw)
        eff = detector_cut(batch, sh, z, method="Efficiency", **kw)
        assert np.allclose(tot, g[f"{tag}/total"], rtol=1e-12)
        assert np.allclose(eff, g[f"{tag}/eff"], rtol=1e-12)
        m = detector_cut(batch, sh, z, method="SampleW", **kw)
        want = g[f"{tag}/mask"]
        if tag == "b":                                   # the reference drops energy-cut particles before masking
            m = m[:, g["b/kept"]]
        assert np.array_equal(m, want)


def test_propagate_particle_api():
    """Shower.propagate_particle (shower.py:509-601): mutates and returns the particle; the result is the oracle's propagation
    with the same Philox key (root key of the shower id the call consumed)."""
    from petite_b200 import Particle
    from petite_b200.constants import m_electron
    from oracle.shower import OracleShower, OParticle
    from oracle.draws import CounterDraws
    from oracle.philox import root_key
    sh = shower("lead", 0.010, seed=21)
    o = OracleShower(None, "lead", 0.010, seed=21, rng="counter")
    for pid, E, m in ((11, 3.0, m_electron), (-11, 0.7, m_electron), (22, 2.0, 0.0)):
        p = Particle([E, 0, 0, np.sqrt(E * E - m * m)], [0.1, 0.2, 0.3], {"PID": pid, "ID": 1, "mass": m})
        fid = sh._next_shower_id
        ret = sh.propagate_particle(p, Losses=(sh._dEdx * 0.1 if pid != 22 else False), MS=(pid != 22))
        assert ret is p and p.get_ended() is True
        q = OParticle(p.get_p0(), p.get_r0(), PID=pid, ID=1, mass=m)
        q.draws = CounterDraws(root_key(21, fid))
        o.propagate(q, o.dEdx * 0.1 if pid != 22 else False, pid != 22)
        assert np.allclose(p.get_pf(), q.pf, rtol=0, atol=1e-9 * E) and np.allclose(p.get_rf(), q.rf, rtol=0, atol=1e-9)
        if pid != 22:
            assert p.get_pf()[0] < E          # energy was lost
    low = Particle([0.001, 0, 0, 0.00085], [0, 0, 0], {"PID": 11, "ID": 1, "mass": m_electron})
    sh.propagate_particle(low, Losses=sh._dEdx * 0.1, MS=True)
    assert np.array_equal(low.get_pf(), low.get_p0()) and low.get_ended()          # below threshold: untouched (shower.py:534-536)
    with pytest.raises(NotImplementedError):
        sh.propagate_particle(Particle([1.0, 0, 0, 1.0], [0, 0, 0], {"PID": 11, "ID": 1, "mass": m_electron}), Losses=False)


def test_dark_sampling_and_produce_bsm_particle_api():
    """DarkShower.draw_dark_sample (dark_shower.py:649-704) and produce_bsm_particle (:721-804) on the Python class."""
    from petite_b200 import Particle
    from petite_b200.constants import m_electron
    from tests.test_gpu_dark import dark_shower
    ds = dark_shower("graphite", 0.03)
    x = ds.draw_dark_sample(5.0, process="DarkBrem")
    assert x.shape == (3,) and 0 < x[0] < 1 and x[1] < 0.31
    xv = ds.draw_dark_sample(5.0, process="DarkBrem", VB=True)
    assert xv.shape == (4,) and xv[3] >= 1
    assert ds.draw_dark_sample(5.0, process="DarkAnn").shape == (1,)
    with pytest.raises(Exception):
        ds.draw_dark_sample(5.0, process="Brem")
    p = Particle([5.0, 0, 0, np.sqrt(25 - m_electron ** 2)], [0, 0, 0.5], {"PID": 11, "ID": 3, "mass": m_electron, "weight": 0.5})
    v = ds.produce_bsm_particle(p, "DarkBrem")
    wg = ds.GetBSMWeights(p, "DarkBrem")
    ids = v.get_ids()
    assert ids["PID"] == 4900022 and ids["parent_ID"] == 3 and ids["ID"] == 6 and ids["generation_process"] == "DarkBrem"
    assert abs(ids["weight"] - 0.5 * wg) <= 1e-12 * wg
    assert 0.03 <= v.get_p0()[0] <= 5.0
    v2 = ds.produce_bsm_particle(p, "DarkBrem", weight=2 * wg)
    assert abs(v2.get_ids()["weight"] - wg) <= 1e-12 * wg
    assert ds.produce_bsm_particle(p, "DarkAnn") is None                          # an electron has no annihilation weight


def test_eta_two_body_bsm_decay():
    """eta / eta' -> gamma V (dark_shower.py:633-638, particle.py meson_decay_dict): weight 2 eps^2 (1 - mV^2/m^2)^3 BR, V on its
    mass shell, energy inside the two-body range of the boosted parent."""
    from petite_b200 import Particle
    from petite_b200.constants import MASS
    from tests.test_gpu_dark import dark_shower
    ds = dark_shower("graphite", 0.03)
    for pid, br in ((221, 0.3936), (331, 0.02307), (111, 0.98823)):
        m = MASS[pid]
        E = 4.0
        p = Particle([E, 0.0, 0.0, np.sqrt(E * E - m * m)], [0, 0, 0], {"PID": pid, "ID": 1, "mass": m, "stability": "short-lived"})
        sm, vs = ds.generate_dark_shower(ExDir=[p])
        assert len(vs) == 1
        v = vs[0]
        w = 2 * ds.kinetic_mixing ** 2 * (1 - (ds._mV / m) ** 2) ** 3 * br
        assert abs(v.get_ids()["weight"] - w) <= 1e-14 * w and v.get_ids()["generation_process"] == "TwoBody_BSMDecay"
        pv = np.asarray(v.get_p0())
        assert abs(pv[0] ** 2 - pv[1:] @ pv[1:] - ds._mV ** 2) < 1e-9
        Ecm = (m * m + ds._mV ** 2) / (2 * m)
        pcm = np.sqrt(Ecm ** 2 - ds._mV ** 2)
        g, b = E / m, np.sqrt(1 - (m / E) ** 2)
        assert g * (Ecm - b * pcm) - 1e-9 <= pv[0] <= g * (Ecm + b * pcm) + 1e-9
