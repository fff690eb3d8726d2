"""VEGAS map training: host refinement step on CPU, full training on the GPU."""
import numpy as np
import pytest

from tests.conftest import DATA


def test_refine_concentrates_increments_where_the_integrand_is():
    from petite_b200.train import refine, integration_range
    g = np.linspace(0.0, 1.0, 101)
    x = 0.5 * (g[1:] + g[:-1])
    f = np.exp(-0.5 * ((x - 0.3) / 0.02) ** 2) + 1e-3
    new = g
    for _ in range(10):                                   # iterate with exact training data for the current grid
        xc = 0.5 * (new[1:] + new[:-1])
        w = np.diff(new) * 100
        fx = np.exp(-0.5 * ((xc - 0.3) / 0.02) ** 2) + 1e-3
        new = refine(new, (w * fx) ** 2, np.ones(100), alpha=0.5)
    assert new[0] == 0.0 and new[-1] == 1.0 and np.all(np.diff(new) > 0)
    inside = np.sum((new > 0.24) & (new < 0.36))
    assert inside > 60                                     # most nodes moved into the +-3 sigma peak region
    # after convergence every increment carries a similar share of the integral
    xc = 0.5 * (new[1:] + new[:-1]); fx = np.exp(-0.5 * ((xc - 0.3) / 0.02) ** 2) + 1e-3
    share = np.diff(new) * fx
    assert share.max() / np.median(share) < 6
    assert integration_range("Brem", 1.0) == [[0, 1], [0, 2], [-2, 2], [0, 1]]
    r = integration_range("MuonBrem", 10.0)
    assert r[0] == [0.001, 10.0 - 0.00051099895] and abs(r[1][1] - np.sqrt(10.0 / 0.00051099895)) < 1e-9     # Q-5 domain


@pytest.mark.gpu
@pytest.mark.parametrize("process", ["PairProd", "Brem", "Comp"])
def test_trained_maps_reproduce_shipped_cross_sections(process):
    """Train on hydrogen as the reference does (default schedule, training weight |jac f|^8), then integrate graphite through the
    NEW maps: sigma must reproduce the shipped sm_xsec rows, and the accept/reject efficiency sigma / (B max_F) must reach the shipped
    maps' (row f-2; measured 2.3-2.8x for the 4-D processes, profiles/r02_final/exp_train_pow2.log).  The plain VEGAS criterion
    (power = 2) is kept as the comparison: same sigma, lower efficiency."""
    from petite_b200.train import Trainer
    from petite_b200 import tables as tb
    from petite_b200.shower import Shower, process_code
    xs = np.load(DATA + "sm_xsec.npz")[f"{process}/graphite"]
    rows = [20, 30, 40, 50, 60, 70, 80, 90, 99]
    E = xs[rows, 0]
    tr = Trainer()
    grids, ninc, I = tr.train(process, E)
    assert np.all(np.isfinite(grids)) and np.all(I > 0)
    off = 0
    for n in ninc:
        assert np.all(np.diff(grids[:, off:off + n + 1], axis=1) > 0)             # nodes stay ordered
        off += n + 1
    sh = Shower(DATA, "graphite", 0.010, seed=3)
    shipped = sh._maps[process]
    mf_old, sg_old = sh.find_max(process, n_trials=2000, seed=9)     # 6e5 points per row: MC error of sigma well below the 5 % bound
    eff_old = (sg_old / (300 * mf_old))[rows]

    def through(g):
        ms = tb.MapSet(process, E, ninc, g, np.ones(len(E)), shipped.neval, shipped.Eg_min, shipped.Ee_min)
        sh._upload_maps(process_code[process], ms)
        sh._maps[process] = ms
        mf, sg = sh.find_max(process, n_trials=2000, seed=9)
        assert np.all(np.abs(sg / xs[rows, 1] - 1) < 0.05), sg / xs[rows, 1]
        return sg / (300 * mf) / eff_old

    r8 = through(grids)
    gmean = float(np.exp(np.mean(np.log(r8))))
    print(process, "efficiency / shipped, power 8:", r8.round(3), "geometric mean", round(gmean, 3))
    if process in ("PairProd", "Brem"):
        # max_F is the maximum of a heavy-tailed sample: a single row can come out several times better or worse than the shipped
        # map's (observed 0.3x-6x, run to run - the fp64 atomics of the training sums are not ordered), so the statement is about the set
        assert gmean > 1.0 and np.median(r8) > 1.0 and np.all(r8 > 0.05), r8
        g2, _, _ = tr.train(process, E, power=2.0)
        r2 = through(g2)
        print(process, "efficiency / shipped, power 2:", r2.round(3))
        assert gmean > float(np.exp(np.mean(np.log(r2))))
    else:
        assert np.all(r8 > 0.12), r8


@pytest.mark.gpu
def test_trained_dark_brem_maps_reproduce_shipped_dark_xsec():
    from petite_b200.train import Trainer
    from petite_b200 import tables as tb
    from petite_b200.dark_shower import DarkShower
    xs = np.load(DATA + "dark_xsec.npz")["0.03/DarkBrem/graphite"]
    rows = [20, 50, 80, 99]
    E = xs[rows, 0]
    tr = Trainer(mT=200.0, mV=0.03)                      # the reference trains on hydrogen with mT = 200 GeV (map readme)
    grids, ninc, _ = tr.train("DarkBrem", E)
    ds = DarkShower(DATA, "graphite", 0.010, 0.03, active_processes=["DarkBrem", "DarkComp", "DarkAnn"], seed=2)
    old = ds._dark_maps["DarkBrem"]
    ms = tb.MapSet("DarkBrem", E, ninc, grids, np.ones(len(E)), old.neval, old.Eg_min, old.Ee_min)
    ds._upload_maps(8, ms)
    ds._dark_maps["DarkBrem"] = ms
    mf, sg = ds.find_max("DarkBrem", n_trials=100, seed=4)
    assert np.all(np.abs(sg / xs[rows, 1] - 1) < 0.05), sg / xs[rows, 1]
    assert np.all(sg / (300 * mf) > 0.01)


@pytest.mark.gpu
def test_retrain_maps_same_showers_statistically_fewer_trials():
    """Shower.retrain_maps (rows f-2 + f-1 on a live engine): accept/reject is exact for any map whose max_F bounds jac f, so the
    showers must be the same PHYSICS (per-shower multiplicity and photon count distributions: two-sample KS; the means within
    4 standard errors) while the sampler spends fewer trials per sample.  4 000 showers of 3 GeV photons in lead each way."""
    from scipy import stats
    import torch
    from petite_b200.shower import Shower
    from tests.gpu_util import primaries

    def run(sh, first):
        b = sh.generate_showers(primaries(22, 3.0, 4000), first_shower_id=first)
        t, m = b._t, b.n
        sid = t["meta"][:m, 3].long()
        mult = torch.bincount(sid, minlength=4000).cpu().numpy()
        phot = torch.bincount(sid, weights=(t["meta"][:m, 0] == 22).double(), minlength=4000).cpu().numpy()
        return mult, phot, b.counters["n_trials"] / b.counters["n_samples"]

    sh = Shower(DATA, "lead", 0.010, seed=31)
    m0, g0, tps0 = run(sh, 0)
    gain = sh.retrain_maps()
    assert set(gain) == {"Brem", "PairProd"}
    for P, (sig, eff) in gain.items():
        ok = np.isfinite(sig) & (sig > 0)
        assert abs(np.median(sig[ok]) - 1) < 0.02, (P, np.median(sig[ok]))           # same cross-sections through the new maps
        assert np.exp(np.mean(np.log(eff[ok & np.isfinite(eff) & (eff > 0)]))) > 1.0, P
    m1, g1, tps1 = run(sh, 100_000)                                                  # other shower ids: independent showers
    print("trials per sample", tps0, "->", tps1, "multiplicity", m0.mean(), m1.mean())
    assert tps1 < 0.9 * tps0
    for a, b in ((m0, m1), (g0, g1)):
        assert stats.ks_2samp(a, b).pvalue > 1e-3
        assert abs(a.mean() - b.mean()) < 4 * np.sqrt(a.var() / len(a) + b.var() / len(b))
