"""Dark pass on the GPU vs the CPU oracle (counter-mode replay) and ensemble checks."""
import numpy as np
import pytest

from tests.conftest import DATA
from tests.gpu_util import primaries

pytestmark = pytest.mark.gpu
_DS = {}
CODE = {"DarkBrem": 8, "DarkAnn_bound": 9, "DarkComp_bound": 10, "DarkMuonBrem": 11, "TwoBody_BSMDecay": 13}


def dark_shower(material, mV, Emin=0.010, seed=31):
    from petite_b200.dark_shower import DarkShower
    key = (material, mV, Emin, seed)
    if key not in _DS:
        _DS[key] = DarkShower(DATA, material, Emin, mV, seed=seed)
    return _DS[key]


def oracle_dark(prims, material, mV, Emin, seed, first_id):
    from oracle.dark import OracleDarkShower
    from oracle.shower import OParticle
    o = OracleDarkShower(None, material, Emin, mV, seed=seed, rng="counter")
    out = []
    for i, p in enumerate(prims):
        ids = p.get_ids()
        sm = o.generate_shower(OParticle(p.get_p0(), p.get_r0(), PID=ids["PID"], ID=ids["ID"], mass=ids["mass"],
                                         stability=ids["stability"]), shower_id=first_id + i)
        out.append(o.generate_dark_shower(sm))
    return out


CASES = [("graphite", 0.03, 11, 5.0, 3), ("graphite", 0.03, 22, 8.0, 2), ("lead", 0.03, -11, 2.0, 4),
         ("lead", 0.03, 13, 6.0, 2), ("graphite", 0.003, 11, 3.0, 3), ("graphite", 1.0, 11, 10.0, 2),
         ("graphite", 0.03, 111, 6.0, 3)]


@pytest.mark.parametrize("material,mV,pid,E,n", CASES)
def test_dark_replay_parity_with_oracle(material, mV, pid, E, n):
    ds = dark_shower(material, mV)
    prims = primaries(pid, E, n, stability="short-lived" if pid == 111 else "stable")
    sm = ds.generate_showers(prims, first_shower_id=7000)
    dk = ds.generate_dark_showers(sm)
    ref = oracle_dark(prims, material, mV, 0.010, 31, 7000)
    h = dk.to_host()
    order, offs = dk.reference_order()
    sm_order, sm_offs = sm.reference_order()
    rank = np.empty(sm.n, dtype=np.int64)
    for i in range(n):
        sl = sm_order[sm_offs[i]:sm_offs[i + 1]]
        rank[sl] = np.arange(len(sl))
    total = 0
    for i, (osm, ovs) in enumerate(ref):
        js = order[offs[i]:offs[i + 1]]
        assert len(js) == len(ovs), (i, len(js), len(ovs))
        assert np.array_equal(h["process"][js], [CODE[v.process] for v in ovs])
        assert np.array_equal(rank[h["parent"][js]], [v.parent_index for v in ovs])
        assert np.array_equal(h["ntrials"][js], [v.ntrials for v in ovs])
        w = np.array([v.weight for v in ovs])
        assert np.all(np.abs(h["weight"][js] - w) <= 1e-9 * np.abs(w) + 1e-300)
        p0 = np.array([v.p0 for v in ovs]).reshape(-1, 4)
        scale = np.maximum(np.max(np.abs(p0), axis=1, keepdims=True), 1e-300)
        assert np.max(np.abs(h["p0"][js] - p0) / scale, initial=0) < 1e-6
        r0 = np.array([v.r0 for v in ovs]).reshape(-1, 3)
        assert np.max(np.abs(h["r0"][js] - r0) / np.maximum(1.0, np.abs(r0)), initial=0) < 1e-6
        total += len(js)
    assert total == dk.n and total > 0


def test_generate_dark_shower_reference_api():
    ds = dark_shower("graphite", 0.03)
    p0 = primaries(11, 3.0, 1)[0]
    sm, vs = ds.generate_dark_shower(SParams=p0)
    assert len(sm) > 10 and len(vs) > 5
    assert all(v.get_ids()["PID"] == 4900022 for v in vs)
    assert all(v.get_ids()["generation_process"] in CODE for v in vs)
    ids = {p.get_ids()["ID"] for p in sm}
    assert all(v.get_ids()["parent_ID"] in ids for v in vs)
    assert abs(vs[0].get_ids()["mass"] - 0.03) < 2e-6                       # Q-21: back-computed, rounded
    sm2, vs2 = ds.generate_dark_shower(ExDir=list(sm))                       # re-processing an existing shower
    assert len(sm2) == len(sm) and len(vs2) > 0
    assert ds.generate_dark_shower() is None


def test_bsm_weights_host_twin_vs_reference_golden(golden):
    g = golden("dark")
    ds = dark_shower("graphite", 0.03)
    E = g["0/w/E"]
    for pid, pr in ((11, "DarkBrem"), (-11, "DarkBrem"), (-11, "DarkAnn"), (22, "DarkComp"), (13, "DarkMuonBrem"), (11, "DarkAnn")):
        want = g[f"0/w/{pid}/{pr}"]
        got = np.array([float(ds.GetBSMWeights([pid, float(e)], pr)) for e in E])
        fin = np.isfinite(want)
        assert np.all(np.abs(got[fin] - want[fin]) <= 1e-12 * np.abs(want[fin])), (pid, pr)


def test_dark_yield_statistics_vs_oracle():
    """Dark-vector yield per primary, weights and energies: GPU ensemble vs an independent oracle sample.  The weight
    distribution is heavy-tailed (resonant annihilation), so distribution-free KS tests are used, not means."""
    ds = dark_shower("graphite", 0.03)
    n_gpu, n_orc = 3000, 60
    prims = primaries(11, 2.0, n_gpu)
    sm = ds.generate_showers(prims, first_shower_id=50_000)
    dk = ds.generate_dark_showers(sm)
    h = dk.to_host()
    y_gpu = np.bincount(h["shower"], weights=h["weight"], minlength=n_gpu)
    ref = oracle_dark(prims[:n_orc], "graphite", 0.03, 0.010, 31, 0)
    y_orc = np.array([sum(v.weight for v in vs) for _, vs in ref])
    from scipy.stats import ks_2samp
    assert ks_2samp(y_gpu, y_orc).pvalue > 0.01
    n_gpu_v = np.bincount(h["shower"], minlength=n_gpu)
    assert ks_2samp(n_gpu_v, np.array([len(vs) for _, vs in ref])).pvalue > 0.01
    # Per-vector quantities are compared through per-shower summaries: the vectors of one shower are correlated, so pooling
    # them would violate the independence the KS test assumes (and a strided sub-sample of the stack would depend on the
    # append order, which is not deterministic); showers are the independent units.
    def per_shower(shower, values, n, fn):
        o = np.argsort(shower, kind="stable")
        cuts = np.searchsorted(shower[o], np.arange(n + 1))
        return np.array([fn(values[o[cuts[i]:cuts[i + 1]]]) for i in range(n) if cuts[i + 1] > cuts[i]])
    lw_gpu = per_shower(h["shower"], np.log10(np.maximum(h["weight"], 1e-300)), n_gpu, np.median)
    lw_orc = np.array([np.median(np.log10(np.maximum([v.weight for v in vs], 1e-300))) for _, vs in ref if vs])
    assert ks_2samp(lw_gpu, lw_orc).pvalue > 0.01
    E_gpu = per_shower(h["shower"], h["p0"][:, 0], n_gpu, np.mean)
    E_orc = np.array([np.mean([v.p0[0] for v in vs]) for _, vs in ref if vs])
    assert ks_2samp(E_gpu, E_orc).pvalue > 0.01
    Emax_gpu = per_shower(h["shower"], h["p0"][:, 0], n_gpu, np.max)
    Emax_orc = np.array([np.max([v.p0[0] for v in vs]) for _, vs in ref if vs])
    assert ks_2samp(Emax_gpu, Emax_orc).pvalue > 0.01


def test_dark_pass_over_concurrent_sub_batches():
    """Parts of run_arrays_split live on other engine handles' stacks; the dark pass over each of them (on the DarkShower's
    own engine, which holds the dark tables) gives the dark vectors of the single-batch run."""
    ds = dark_shower("graphite", 0.03)
    prims = primaries(11, 3.0, 24)
    one = ds.generate_showers(prims, first_shower_id=9100)
    d_one = ds.generate_dark_showers(one)
    h = d_one.to_host()
    ref_n, ref_w, ref_E = d_one.n, np.sort(h["weight"]), np.sort(h["p0"][:, 0])
    parts = ds.generate_showers_split(prims, parts=2, first_shower_id=9100)
    assert parts[0]._owner is not parts[1]._owner and sum(p.n for p in parts) == one.n
    ws, Es, n = [], [], 0
    for part in parts:
        d = ds.generate_dark_showers(part)
        hh = d.to_host()
        n += d.n; ws.append(hh["weight"].copy()); Es.append(hh["p0"][:, 0].copy())
    assert n == ref_n
    assert np.array_equal(np.sort(np.concatenate(ws)), ref_w)
    assert np.array_equal(np.sort(np.concatenate(Es)), ref_E)
