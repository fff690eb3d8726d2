"""Oracle dark pass vs the UNMODIFIED reference DarkShower (stream mode, tests/golden/dark.npz)."""
import random

import numpy as np
import pytest

from oracle.consts import m_electron, m_muon, m_pi0
from oracle.dark import OracleDarkShower
from oracle.shower import OParticle

CODE = {"DarkBrem": 8, "DarkAnn_bound": 9, "DarkComp_bound": 10, "DarkMuonBrem": 11, "TwoBody_BSMDecay": 13}
CASES = [(ci, k) for ci in range(4) for k in range(5)]
_ORC = {}


def orc(material, mV):
    if (material, mV) not in _ORC:
        _ORC[(material, mV)] = OracleDarkShower(None, material, 0.010, mV, rng="stream")
    return _ORC[(material, mV)]


@pytest.mark.parametrize("ci", range(4))
def test_bsm_weights(golden, ci):
    g = golden("dark")
    o = orc(str(g[f"{ci}/material"]), float(g[f"{ci}/mV"]))
    E = g[f"{ci}/w/E"]
    mass = {11: m_electron, -11: m_electron, 22: 0.0, 13: m_muon}
    for pid, pr in ((11, "DarkBrem"), (-11, "DarkBrem"), (-11, "DarkAnn"), (22, "DarkComp"), (13, "DarkMuonBrem"), (11, "DarkAnn")):
        want = g[f"{ci}/w/{pid}/{pr}"]
        got = np.array([o.bsm_weight(pid, float(e), mass[pid], pr) for e in E])
        fin = np.isfinite(want)
        assert np.array_equal(np.isfinite(got), fin)          # x/0 -> inf above the table range, as in the reference
        assert np.all(np.abs(got[fin] - want[fin]) <= 1e-12 * np.abs(want[fin])), (pid, pr)
    assert abs(o.bsm_weight(111, 5.0, m_pi0, "TwoBody_BSMDecay") - g[f"{ci}/w/pi0"][0]) <= 1e-15


@pytest.mark.parametrize("ci,k", CASES)
def test_dark_shower_stream_mode_equals_reference(golden, ci, k):
    g = golden("dark")
    material, mV = str(g[f"{ci}/material"]), float(g[f"{ci}/mV"])
    o = orc(material, mV)
    pre = f"{ci}/sh{k}/"
    pid, E0, seed, n_sm = g[pre + "case"]
    pid, seed = int(pid), int(seed)
    m = {11: m_electron, -11: m_electron, 22: 0.0, 13: m_muon, 111: m_pi0}[pid]
    np.random.seed(seed)
    random.seed(seed)
    sm = o.generate_shower(OParticle([E0, 0, 0, np.sqrt(E0 ** 2 - m ** 2)], [0, 0, 0], PID=pid, ID=1, mass=m,
                                     stability="short-lived" if pid == 111 else "stable"))
    assert len(sm) == int(n_sm)
    _, vs = o.generate_dark_shower(sm)
    assert len(vs) == len(g[pre + "weight"])
    assert np.array_equal([CODE[v.process] for v in vs], g[pre + "process"])
    assert np.array_equal([v.parent_PID for v in vs], g[pre + "parent_PID"])
    assert np.array_equal([v.parent_ID % (1 << 61) for v in vs], g[pre + "parent_ID_mod"])
    w = np.array([v.weight for v in vs])
    assert np.all(np.abs(w - g[pre + "weight"]) <= 1e-11 * np.abs(g[pre + "weight"]))
    p0 = np.array([v.p0 for v in vs]).reshape(-1, 4)
    scale = np.maximum(np.max(np.abs(g[pre + "p0"]), axis=1, keepdims=True), 1e-300)
    assert np.max(np.abs(p0 - g[pre + "p0"]) / scale, initial=0) < 1e-6
    r0 = np.array([v.r0 for v in vs]).reshape(-1, 3)
    rscale = np.maximum(np.max(np.abs(g[pre + "r0"]), axis=1, keepdims=True), 1e-3)
    assert np.max(np.abs(r0 - g[pre + "r0"]) / rscale, initial=0) < 1e-6
    assert np.allclose([v.mass for v in vs], g[pre + "mass"], rtol=0, atol=1.1e-6)      # Q-21: rounded to 6 decimals
