"""Statistical parity at BASELINE size: >= 1e5 GPU showers per configuration against an oracle ensemble
(tests/golden/ensemble.npz, made by tests/golden/make_ensemble.py).  KS tests on per-shower observables (multiplicity,
photon/positron counts, energy, depth, lateral spread and angle summaries, dark-vector yield and weights) and a chi-square
on the per-shower photon energy spectrum, all at p > 0.01 (BASELINE.json north_star)."""
import numpy as np
import pytest

from tests.conftest import DATA
from tests import ensemble_stats as es

pytestmark = pytest.mark.gpu
N_GPU = 100_000
SM_KEYS = ["mult", "n_gamma", "n_eplus", "E_gamma", "Emax_sec", "z_mean", "rT_mean", "theta_e"]
DARK_KEYS = ["n_V", "dyield", "lw_med", "EV_mean", "EV_max"]


def _engine(name):
    cfg = es.CONFIGS[name]
    if cfg["mV"] is None:
        from petite_b200.shower import Shower
        return Shower(DATA, cfg["material"], cfg["E_min"], seed=cfg["seed"])
    from petite_b200.dark_shower import DarkShower
    return DarkShower(DATA, cfg["material"], cfg["E_min"], cfg["mV"], seed=cfg["seed"])


def _run(sh, name, n, first_id=0):
    cfg = es.CONFIGS[name]
    E, m = cfg["E0"], cfg["mass"]
    p = np.tile([E, 0.0, 0.0, np.sqrt(E * E - m * m)], (n, 1))
    arrays = (p, np.zeros((n, 3)), np.ones(n), np.full(n, m), np.full(n, cfg["pid"], dtype=np.int32), np.zeros(n, dtype=np.int32))
    return sh.run_arrays(*arrays, first_shower_id=first_id)


def _compare(gpu, golden, name, keys):
    from scipy.stats import ks_2samp
    pvals = {}
    for k in keys:
        pvals[k] = ks_2samp(gpu[k], golden[f"{name}/{k}"]).pvalue
    return pvals


def _spectrum_chi2(gpu_spec, orc_spec):
    """Per-shower photon counts in 8 energy bins: chi-square of the difference of the ensemble means, each bin's variance
    estimated from the shower-to-shower scatter on both sides."""
    from scipy.stats import chi2
    mg, mo = gpu_spec.mean(0), orc_spec.mean(0)
    var = gpu_spec.var(0, ddof=1) / len(gpu_spec) + orc_spec.var(0, ddof=1) / len(orc_spec)
    use = var > 0
    x2 = float(np.sum((mg[use] - mo[use]) ** 2 / var[use]))
    return x2, int(use.sum()), float(chi2.sf(x2, int(use.sum())))


# showers per batch: the 100 GeV muon showers keep 7.9e3 records each, so 1e5 of them are stepped as four batches
BATCH = {"c5_mu_lead": 25_000}


@pytest.mark.parametrize("name", ["c2_gamma_lead", "c1_e_graphite", "c5_mu_lead"])
def test_sm_observables_1e5_showers(name, golden):
    g = golden("ensemble")
    sh = _engine(name)
    nb = BATCH.get(name, N_GPU)
    parts = []
    for first in range(0, N_GPU, nb):
        batch = _run(sh, name, nb, first_id=first)
        assert batch.counters["n_no_sample"] == 0
        parts.append(es.summarise_gpu_sm(batch, nb))
        del batch
    gpu = {k: np.concatenate([q[k] for q in parts]) for k in parts[0]}
    assert len(gpu["mult"]) == N_GPU
    pvals = _compare(gpu, g, name, SM_KEYS)
    x2, ndf, p_spec = _spectrum_chi2(gpu["spec"], g[f"{name}/spec"])
    print(name, {k: round(float(v), 4) for k, v in pvals.items()}, "spectrum chi2/ndf", round(x2, 2), ndf, "p", round(p_spec, 4))
    assert all(v > 0.01 for v in pvals.values()), pvals
    assert p_spec > 0.01, (x2, ndf, p_spec)
    # energy bookkeeping at full size (size-independent property): no secondary is created above the primary's energy
    assert np.all(gpu["Emax_sec"] <= es.CONFIGS[name]["E0"] * (1 + 1e-12))
    del sh


def test_dark_observables_1e5_showers(golden):
    name = "c3_dark_graphite"
    g = golden("ensemble")
    sh = _engine(name)
    batch = _run(sh, name, N_GPU)
    dk = sh.generate_dark_showers(batch)
    gpu = es.summarise_gpu_sm(batch, N_GPU)
    gpu.update(es.summarise_gpu_dark(dk, N_GPU))
    pvals = _compare(gpu, g, name, ["mult", "E_gamma", "z_mean"] + DARK_KEYS)
    print(name, {k: round(float(v), 4) for k, v in pvals.items()}, "dark vectors", dk.n)
    assert all(v > 0.01 for v in pvals.values()), pvals
    assert dk.n > 50 * N_GPU


def test_reference_recorded_single_shower_numbers():
    """The only numbers the reference records for whole showers (stored notebook outputs, single seeds):
    10 GeV e- into graphite, E_min = 10 MeV -> 655 particles (multiple_coulomb_scattering.ipynb:[11]); the profiled shower of
    tutorial.ipynb:[37] made 700 propagate_particle calls and 4 998 get_scattered_momentum_fast calls.  They must be
    ordinary members of the GPU ensemble: inside its central 99 %, and the multiple-scattering calls per step within 15 %."""
    name = "c1_e_graphite"
    sh = _engine(name)
    n = 20_000
    batch = _run(sh, name, n)
    mult = es.summarise_gpu_sm(batch, n)["mult"]
    lo, hi = np.quantile(mult, [0.005, 0.995])
    assert lo <= 655 <= hi and lo <= 700 <= hi, (lo, hi)
    c = batch.counters
    mcs_calls_per_step = (c["n_substeps"] + c["n_charged"]) / c["n_steps"]      # one call per sub-step + one for the final step
    assert abs(mcs_calls_per_step / (4998 / 700) - 1) < 0.15, mcs_calls_per_step
