"""The CPU oracle against golden vectors produced by the UNMODIFIED reference (tests/golden/make_golden.py)."""
import random

import numpy as np
import pytest

from oracle import integrands as I
from oracle import physics as phy
from oracle import philox as ph
from oracle.consts import TARGETS, SM_PROCESSES, DARK_PROCESSES, m_electron, m_muon, m_pi0, PROC_CODE
from oracle.shower import OracleShower, OParticle

RTOL = 1e-12   # north_star: deterministic pieces within 1e-12 relative


def test_philox_known_answers():
    # Random123 kat_vectors for philox4x32-10
    kat = [((0, 0, 0, 0), (0, 0), "6627e8d5 e169c58d bc57ac4c 9b00dbd8"),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, "408f276d 41c83b0e a20bc7c6 6d5451fd"),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), "d16cfe09 94fdcceb 5001e420 24126ea1")]
    for c, k, want in kat:
        got = " ".join("%08x" % int(v) for v in ph.philox4x32(*c, *k))
        assert got == want


@pytest.mark.parametrize("material", ["graphite", "lead"])
@pytest.mark.parametrize("process", SM_PROCESSES + DARK_PROCESSES)
def test_integrands(golden, material, process):
    g = golden("integrands")
    E, x, f = g[f"{material}/{process}/E"], g[f"{material}/{process}/x"], g[f"{material}/{process}/f"]
    t = TARGETS[material]
    dim = {"Brem": 4, "PairProd": 4, "MuonBrem": 4, "DarkBrem": 3, "DarkMuonBrem": 3}.get(process, 1)
    for Einc in np.unique(E):
        sel = E == Einc
        ev = dict(E_inc=float(Einc), Z_T=t["Z_T"], A_T=t["A_T"], mT=t["A_T"], mV=0.0 if process in SM_PROCESSES else 0.03,
                  Eg_min=0.001, Ee_min=0.005, m_lepton=m_muon if "Muon" in process else m_electron)
        got = I.DSIGMA[process](x[sel][:, :dim], ev)
        want = f[sel]
        scale = np.max(np.abs(want)) if np.any(want != 0) else 1.0
        # 1e-12 relative; points sitting on a cancellation (dark brem W -> 0) get a mixed abs/rel bound
        tol = RTOL * np.abs(want) + (1e-9 * scale if "Dark" in process and "Brem" in process else 0.0)
        assert np.all(np.abs(got - want) <= tol + 1e-300), (process, Einc, np.max(np.abs(got - want) / (np.abs(want) + 1e-300)))
        assert np.array_equal(want == 0, got == 0)


def test_kinematics(golden):
    g = golden("kinematics")
    fn = {"Brem": lambda a: phy.kin_brem(a[0], a[1], a[2:6], a[6]), "MuonBrem": lambda a: phy.kin_brem(a[0], a[1], a[2:6], a[6]),
          "PairProd": lambda a: phy.kin_pairprod(a[0], a[2:6], a[6]), "Comp": lambda a: phy.kin_compton(a[0], a[2:6], a[6]),
          "Ann": lambda a: phy.kin_annihilation(a[0], a[2:6], a[6]), "Moller": lambda a: phy.kin_ee(a[0], a[2:6], a[6]),
          "Bhabha": lambda a: phy.kin_ee(a[0], a[2:6], a[6]), "MuonE": lambda a: phy.kin_mue(a[0], a[2:6], a[6]),
          "SMDecay": lambda a: phy.two_body_decay([a[0], a[2], a[3], a[4]], a[1], 0.0, 0.0, a[6], a[7])}
    for P, f in fn.items():
        for a, want in zip(g[f"{P}/in"], g[f"{P}/out"]):
            v1, v2 = f(a)
            got = np.array(list(v1) + list(v2))
            scale = np.max(np.abs(want))
            assert np.all(np.abs(got - want) <= 1e-12 * scale), (P, a, got, want)


def test_multiple_scattering(golden):
    g = golden("mcs")
    for a, want in zip(g["inp"], g["out"]):
        mat = "graphite" if a[10] == 6 else "lead"
        t = TARGETS[mat]
        got = phy.mcs_scatter(list(a[:4]), t["rho"] * (a[4] / 0.01), t["A_T"], t["Z_T"], 1, a[5], a[6], a[7], a[8], a[9])
        assert np.all(np.abs(np.array(got) - want) <= 1e-12 * np.max(np.abs(want))), (a, got, want)


def test_particle_pieces(golden):
    g = golden("particle")
    for a, want in zip(g["lose_in"], g["lose_out"]):
        got = phy.lose_energy(list(a[:4]), a[4], a[5])
        assert np.allclose(got, want, rtol=1e-13, atol=0)
    for a, want in zip(g["rot_in"], g["rot_out"]):
        got = np.array(phy.rotation_matrix(list(a))).ravel()
        assert np.allclose(got, want, rtol=0, atol=1e-14)


@pytest.mark.parametrize("material", ["graphite", "lead"])
def test_nsigma_and_mfp(golden, material):
    g = golden("nsigma")
    o = OracleShower(None, material, 0.010)
    E = g[f"{material}/E"]
    for P in SM_PROCESSES:
        want = g[f"{material}/{P}"]
        got = np.array([o.NSigma[P](float(e)) for e in E])
        assert np.all(np.abs(got - want) <= 1e-12 * np.abs(want)), P
        assert np.allclose(o.NSigma[P].x, g[f"{material}/{P}/table_x"], rtol=1e-15)
        assert np.allclose(o.NSigma[P].y, g[f"{material}/{P}/table_y"], rtol=1e-13)
    for pid in (22, 11, -11, 13):
        want = g[f"{material}/mfp/{pid}"]
        got = np.array([o.get_mfp(pid, float(e)) for e in E])
        assert np.all(np.abs(got - want) <= 1e-12 * np.abs(want))
    assert np.allclose([o.min_calc[k] for k in (11, -11, 22, 13, -13)], g[f"{material}/min_calc"], rtol=0, atol=0)
    assert np.allclose([o.nT, o.ne], g[f"{material}/n"], rtol=1e-15)


def test_survey_golden_mfp_values():
    """values computed from data/sm_xsec.pkl with the reference's constants."""
    o = OracleShower(None, "graphite", 0.010)
    assert abs(o.NSigma["PairProd"](1.0) / 3.083710e-02 - 1) < 1e-6
    assert abs(o.NSigma["Brem"](10.0) / 4.976817e-01 - 1) < 1e-6
    assert abs(o.get_mfp(22, 10.0) / 3.077685e-01 - 1) < 1e-6
    o = OracleShower(None, "lead", 0.010)
    assert abs(o.NSigma["Comp"](0.01) / 1.373309e-01 - 1) < 1e-6
    assert abs(o.get_mfp(22, 1.0) / 7.009224e-03 - 1) < 1e-6


@pytest.mark.parametrize("case", range(8))
def test_stream_mode_shower_equals_reference(golden, case):
    """Whole generate_shower runs: the oracle fed the reference's own uniform streams must make the same decisions."""
    g = golden("showers")
    pid, E, Emin, seed, mass = g[f"{case}/case"]
    pid, seed = int(pid), int(seed)
    mat = str(g[f"{case}/material"])
    m = {11: m_electron, -11: m_electron, 22: 0.0, 13: m_muon, -13: m_muon, 111: m_pi0}[pid]
    o = OracleShower(None, mat, float(Emin), rng="stream")
    np.random.seed(seed)
    random.seed(seed)
    p0 = OParticle([E, 0, 0, np.sqrt(E ** 2 - m ** 2)], [0, 0, 0], PID=pid, ID=1, mass=None if mass < 0 else float(mass),
                   stability="short-lived" if pid == 111 else "stable")
    got = o.generate_shower(p0)
    assert len(got) == len(g[f"{case}/pid"])
    assert np.array_equal([q.PID for q in got], g[f"{case}/pid"])
    assert np.array_equal([PROC_CODE[q.process] for q in got], g[f"{case}/process"])
    assert np.array_equal([q.gen for q in got], g[f"{case}/gen"])
    assert np.array_equal([q.ID % (1 << 61) for q in got], g[f"{case}/ID_mod"])
    assert np.allclose([q.weight for q in got], g[f"{case}/weight"], rtol=1e-15)
    assert np.allclose([q.mass for q in got], g[f"{case}/mass"], rtol=0, atol=0)
    for name, arr in (("p0", [q.p0 for q in got]), ("pf", [q.pf for q in got]), ("r0", [q.r0 for q in got]), ("rf", [q.rf for q in got])):
        want = g[f"{case}/{name}"]
        arr = np.asarray(arr)
        scale = np.maximum(np.max(np.abs(want), axis=1, keepdims=True), 1e-300)
        # ulp-level differences are amplified by the reference's own ill-conditioned acos(pz/|p|) (particle.py:181)
        assert np.max(np.abs(arr - want) / scale) < 1e-6, name


@pytest.mark.parametrize("case", range(6))
def test_stream_mode_long_lived_shower_equals_reference(golden, case):
    """Decay in flight of pi+- / K+- primaries (particle.py:363-389, 410-422; row f-5): whole generate_shower runs of the
    unmodified reference (tests/golden/make_longlived.py) - decay point, the two independently drawn daughter weights, the
    muon's shower - reproduced by the oracle on the reference's own uniform streams."""
    from oracle import consts as OC
    g = golden("longlived")
    pid, E, Emin, seed, mass = g[f"{case}/case"]
    pid, seed = int(pid), int(seed)
    o = OracleShower(None, str(g[f"{case}/material"]), float(Emin), rng="stream")
    np.random.seed(seed)
    random.seed(seed)
    assert mass == OC.MASS[pid]
    p0 = OParticle([E, 0, 0, np.sqrt(E ** 2 - mass ** 2)], [0, 0, 0], PID=pid, ID=1, mass=float(mass), stability="long-lived")
    got = o.generate_shower(p0)
    assert len(got) == len(g[f"{case}/pid"]) and len(got) > 50
    assert np.array_equal([q.PID for q in got], g[f"{case}/pid"])
    assert got[1].PID == (-13 if pid > 0 else 13) and got[2].PID == (14 if pid > 0 else -14)
    assert np.array_equal([PROC_CODE[q.process] for q in got], g[f"{case}/process"])
    assert np.array_equal([q.gen for q in got], g[f"{case}/gen"])
    assert np.array_equal([q.ID % (1 << 61) for q in got], g[f"{case}/ID_mod"])
    w = np.array([q.weight for q in got])
    assert np.allclose(w, g[f"{case}/weight"], rtol=1e-14) and w[1] != w[2]          # two draws of prob_decay_b_int
    assert np.allclose([q.mass for q in got], g[f"{case}/mass"], rtol=0, atol=0)
    assert np.allclose(got[0].rf, g[f"{case}/rf"][0], rtol=1e-15, atol=0) and got[0].rf[2] > 0       # the decay point
    for name, arr in (("p0", [q.p0 for q in got]), ("pf", [q.pf for q in got]), ("r0", [q.r0 for q in got]), ("rf", [q.rf for q in got])):
        want = g[f"{case}/{name}"]
        arr = np.asarray(arr)
        scale = np.maximum(np.max(np.abs(want), axis=1, keepdims=True), 1e-300)
        assert np.max(np.abs(arr - want) / scale) < 1e-6, name


def test_dark_kinematics(golden):
    """oracle.physics kin_darkbrem / kin_darkann / kin_compton_bound against l_to_lV_fourvecs, radiative_return_fourvecs and
    compton_fourvecs_boundelectron of the unmodified reference (tests/golden/make_golden.py golden_dark_kinematics)."""
    g = golden("dark_kinematics")
    n = 0
    for key in [k for k in g.files if k.endswith("/in")]:
        tag, P, _ = key.split("/")
        for a, want in zip(g[key], g[f"{tag}/{P}/out"]):
            E, mV, x, u1, u2, Pe, cte = a[0], a[1], a[2:6], a[6], a[7], a[8], a[9]
            if P in ("DarkBrem", "DarkMuonBrem"):
                v = phy.kin_darkbrem(E, m_muon if "Muon" in P else m_electron, x, u1, mV)[1]
            elif P == "DarkAnn":
                v = phy.kin_darkann(E, x, mV)[1]
            else:
                v = phy.kin_compton_bound(E, x, mV, Pe, cte, u1, u2)[1]
            assert np.all(np.abs(np.array(v) - want) <= 1e-12 * np.max(np.abs(want))), (P, a, v, want)
            n += 1
    assert n > 600
