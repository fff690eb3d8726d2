"""Statistical parity at BASELINE size: >= 1e5 GPU showers per configuration against
  * ensembles made by the UNMODIFIED reference in stream mode (tests/golden/ensemble_ref.npz, make_ensemble_ref.py; 500 showers
    for configs 1 and 2, 200 for configs 3 and 5 including their dark passes), and
  * the oracle's counter-mode ensembles (tests/golden/ensemble.npz, make_ensemble.py).
KS tests on per-shower observables (multiplicity, photon/positron counts, energy, depth, lateral spread and angle summaries,
dark-vector yield and weights) and a chi-square on the per-shower photon energy spectrum, all at p > 0.01 (BASELINE.json
north_star)."""
import numpy as np
import pytest

from tests.conftest import DATA
from tests import ensemble_stats as es

pytestmark = pytest.mark.gpu
N_GPU = 100_000
SM_KEYS = ["mult", "n_gamma", "n_eplus", "E_gamma", "Emax_sec", "z_mean", "rT_mean", "theta_e"]
DARK_KEYS = ["n_V", "dyield", "lw_med", "EV_mean", "EV_max"]


def _engine(name):
    import os
    from tests.conftest import ROOT
    cfg = es.REF_CONFIGS[name]
    data = os.path.join(ROOT, cfg["data"], "") if cfg.get("data") else DATA
    if cfg["mV"] is None:
        from petite_b200.shower import Shower
        return Shower(data, cfg["material"], cfg["E_min"], seed=cfg["seed"])
    from petite_b200.dark_shower import DarkShower
    return DarkShower(data, cfg["material"], cfg["E_min"], cfg["mV"], seed=cfg["seed"], active_processes=cfg.get("active"))


_PRIM = {}


def _run(sh, name, n, first_id=0):
    """Showers first_id .. first_id + n - 1 of the configuration (primary i is always the same particle, whatever the batching)."""
    cfg = es.REF_CONFIGS[name]
    if cfg["pid"] == 0:
        if name not in _PRIM:
            _PRIM[name] = es.config_primaries(cfg, N_GPU)
        p, pid, m = (a[first_id:first_id + n] for a in _PRIM[name])
    else:
        p, pid, m = es.config_primaries(cfg, n)
    arrays = (np.ascontiguousarray(p), np.zeros((n, 3)), np.ones(n), np.ascontiguousarray(m), np.ascontiguousarray(pid), np.zeros(n, dtype=np.int32))
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


# showers per batch: the 100 GeV muon showers keep 7.9e3 records (+ 6.6e3 dark vectors) each, so 1e5 of them are stepped as
# four batches
BATCH = {"c5_mu_lead_dark": 10_000, "c4_beamdump_lead_dark": 10_000}
# oracle-side (counter mode) twin of a reference-side configuration and the observables it holds
ORACLE_TWIN = {"c2_gamma_lead": ("c2_gamma_lead", SM_KEYS), "c1_e_graphite": ("c1_e_graphite", SM_KEYS),
               "c3_dark_graphite": ("c3_dark_graphite", ["mult", "E_gamma", "z_mean"] + DARK_KEYS), "c5_mu_lead_dark": ("c5_mu_lead", SM_KEYS)}


@pytest.mark.parametrize("name", list(es.REF_CONFIGS))
def test_observables_1e5_showers_vs_reference_and_oracle(name, golden):
    """All five BASELINE configurations (config 4 on the retrained 400 GeV table set, data_400GeV/): the GPU ensemble against the
    reference's own stream-mode ensemble AND (configs 1, 2, 3, 5) the oracle's counter-mode one."""
    ref, orc = golden("ensemble_ref"), golden("ensemble")
    cfg = es.REF_CONFIGS[name]
    dark = cfg["mV"] is not None
    sh = _engine(name)
    nb = BATCH.get(name, N_GPU)
    parts, n_dark = [], 0
    for first in range(0, N_GPU, nb):
        batch = _run(sh, name, nb, first_id=first)
        assert batch.counters["n_no_sample"] == 0
        row = es.summarise_gpu_sm(batch, nb)
        if dark:
            dk = sh.generate_dark_showers(batch)
            row.update(es.summarise_gpu_dark(dk, nb))
            n_dark += dk.n
            del dk
        parts.append(row)
        del batch
    gpu = {k: np.concatenate([q[k] for q in parts]) for k in parts[0]}
    assert len(gpu["mult"]) == N_GPU
    keys = SM_KEYS + (DARK_KEYS if dark else [])
    p_ref = _compare(gpu, ref, name, keys)
    x2, ndf, ps_ref = _spectrum_chi2(gpu["spec"], ref[f"{name}/spec"])
    if name in ORACLE_TWIN:
        twin, okeys = ORACLE_TWIN[name]
        p_orc = _compare(gpu, orc, twin, okeys)
        _, _, ps_orc = _spectrum_chi2(gpu["spec"], orc[f"{twin}/spec"])
    else:                         # config 4: reference-made ensemble only (a 400 GeV oracle ensemble would take hours of CPU)
        p_orc, ps_orc = {}, 1.0
    print(name, "vs reference (stream)", {k: round(float(v), 4) for k, v in p_ref.items()}, "spectrum p", round(ps_ref, 4),
          "| vs oracle (counter)", {k: round(float(v), 4) for k, v in p_orc.items()}, "spectrum p", round(ps_orc, 4), "| dark vectors", n_dark)
    assert all(v > 0.01 for v in p_ref.values()), p_ref
    assert ps_ref > 0.01, (x2, ndf, ps_ref)
    assert all(v > 0.01 for v in p_orc.values()), p_orc
    assert ps_orc > 0.01
    # energy bookkeeping at full size (size-independent property): no secondary is created above the primary's energy plus the rest
    # energy of the atomic electron it may have struck (Compton, Moller / Bhabha, annihilation in flight)
    E_prim = _PRIM[name][0][:N_GPU, 0] if cfg["pid"] == 0 else cfg["E0"]
    assert np.all(gpu["Emax_sec"] <= (E_prim + es.m_e) * (1 + 1e-12))
    if dark:
        assert n_dark > 50 * N_GPU
    del sh


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
