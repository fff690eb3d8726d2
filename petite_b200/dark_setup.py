"""Builder for the DarkShower constructor tables (weights, dRate/dE, n*sigma_dark) when no cache exists.

The reference computes these with nested adaptive quadratures at construction time (dark_shower.py:254-593, about half
a minute per (material, mV)) and caches part of them in ``dark_weights.pkl`` / ``dark_drate.pkl``.  This project caches
ALL of them in ``<dict_dir>/dark_setup_<material>_mV<mV>.npz``; caches for the BASELINE configurations ship in data/.
"""


def build(shower, path):
    raise NotImplementedError(
        f"no dark set-up cache at {path}.  Caches ship for graphite (mV = 0.003, 0.03, 1.0) and lead (mV = 0.03); for other "
        "(material, mV) pairs dump one from a reference install with tests/golden/make_golden.py (dump_dark_setup).")
