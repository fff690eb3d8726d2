"""Counter mode vs stream mode, on the CPU: the oracle's counter-mode ensembles (tests/golden/ensemble.npz, Philox draws keyed by
particle - what the CUDA engine implements, including the shortcuts of oracle/draws.py: one Philox call per multiple-scattering
block with the Box-Muller angle dropped, accept uniforms pre-drawn per trial) against ensembles made by the UNMODIFIED reference
in stream mode (tests/golden/ensemble_ref.npz, make_ensemble_ref.py: numpy.random / random global streams in the reference's
own call order).  Both are samples of the same physics iff the shortcuts are distribution-preserving: two-sample KS on every
per-shower observable and a chi-square on the per-shower photon spectrum, p > 0.01 (BASELINE.json north_star)."""
import os

import numpy as np
import pytest

from tests import ensemble_stats as es
from tests.conftest import GOLDEN

SM_KEYS = ["mult", "n_gamma", "n_eplus", "E_gamma", "Emax_sec", "z_mean", "rT_mean", "theta_e"]
DARK_KEYS = ["n_V", "dyield", "lw_med", "EV_mean", "EV_max"]
# (reference-side configuration, oracle-side configuration, observables)
PAIRS = [("c2_gamma_lead", "c2_gamma_lead", SM_KEYS), ("c1_e_graphite", "c1_e_graphite", SM_KEYS),
         ("c3_dark_graphite", "c3_dark_graphite", SM_KEYS + DARK_KEYS), ("c5_mu_lead_dark", "c5_mu_lead", SM_KEYS)]


def spectrum_chi2(a, b):
    from scipy.stats import chi2
    ma, mb = a.mean(0), b.mean(0)
    var = a.var(0, ddof=1) / len(a) + b.var(0, ddof=1) / len(b)
    use = var > 0
    x2 = float(np.sum((ma[use] - mb[use]) ** 2 / var[use]))
    return x2, int(use.sum()), float(chi2.sf(x2, int(use.sum())))


@pytest.mark.parametrize("ref_name,orc_name,keys", PAIRS)
def test_counter_mode_vs_reference_stream_mode(golden, ref_name, orc_name, keys):
    from scipy.stats import ks_2samp
    ref, orc = golden("ensemble_ref"), golden("ensemble")
    assert len(ref[f"{ref_name}/mult"]) >= es.REF_CONFIGS[ref_name]["n_ref"]
    p = {k: float(ks_2samp(ref[f"{ref_name}/{k}"], orc[f"{orc_name}/{k}"]).pvalue) for k in keys}
    x2, ndf, ps = spectrum_chi2(ref[f"{ref_name}/spec"], orc[f"{orc_name}/spec"])
    print(ref_name, {k: round(v, 4) for k, v in p.items()}, "spectrum chi2/ndf", round(x2, 2), ndf, "p", round(ps, 4))
    assert all(v > 0.01 for v in p.values()), p
    assert ps > 0.01, (x2, ndf, ps)


def test_reference_ensemble_is_reference_made():
    """The fixture carries the reference's wall time per shower (stream mode runs at the reference's speed, seconds per
    shower, not the oracle's vectorised sweeps) and one row per seeded shower."""
    ref = np.load(os.path.join(GOLDEN, "ensemble_ref.npz"))
    for name, cfg in es.REF_CONFIGS.items():
        assert len(ref[f"{name}/seconds"]) == cfg["n_ref"]
        assert np.all(ref[f"{name}/seconds"] > 0)
