"""Closed-form total cross-sections used once, at construction, to tabulate n*sigma(E) (host setup).

Moller, Bhabha and muon-electron scattering with a cut ``Ee_min`` on the struck electron's energy; results equal
the reference's ``sigma_moller`` / ``sigma_bhabha`` / ``sigma_muone`` (all_processes.py:888-1051), which
``Shower.set_NSigmas`` (shower.py:285-293) evaluates on a geometric energy grid.
"""
import numpy as np

from .constants import alpha_em, m_electron as me, m_muon as mmu


def _step(x):
    return np.heaviside(x, 1)


def sigma_moller(E, Ee_min):
    E = np.asarray(E, dtype=np.float64)
    T = Ee_min - me
    thr = 3 * me + 4 * T
    on, off = _step(E - thr), _step(thr - E)
    pref = 2 * np.pi * alpha_em ** 2 / (me * (E ** 2 - me ** 2))
    rational = E - 3 * me - 4 * T + 2 * E ** 2 * (-2 / (E - 3 * me - 2 * T) + 1 / T + 1 / (-E + me + T) + 2 / (E + me + 2 * T))
    with np.errstate(all="ignore"):
        logarg = ((-E + me + T) * (-E + 3 * me + 2 * T) / (T * (E + me + 2 * T))) * on + off
        logterm = 2 * me * (me - 2 * E) / (E - me) * np.log(logarg)
    return pref * (rational + logterm) * on


def sigma_bhabha(E, Ee_min):
    E = np.asarray(E, dtype=np.float64)
    T = Ee_min - me
    thr = 3 * me + 4 * T
    on, off = _step(E - thr), _step(thr - E)
    gap = E - 3 * me - 2 * T
    pref = np.pi * alpha_em ** 2 / (12 * (E - me) * me * (E + me) ** 3 * gap * T)
    poly = (E - 3 * me - 4 * T) * (24 * E ** 2 * (E + me) ** 2
                                   + (E - 3 * me) * (31 * E ** 2 + 84 * E * me + 57 * me ** 2) * T
                                   - 4 * (16 * E ** 2 + 39 * E * me + 33 * me ** 2) * T ** 2
                                   + 8 * (E - 3 * me) * T ** 3 - 8 * T ** 4)
    with np.errstate(all="ignore"):
        logterm = 24 * (E + me) * (2 * E ** 2 + 4 * E * me + me ** 2) * gap * T * np.log((2 * T / gap) * on + off)
    return pref * (poly + logterm) * on


def muone_threshold(Ee_min):
    """Muon energy above which the struck electron can exceed Ee_min (shower.py:289)."""
    return 1.0 / (2.0 * me) * (me * (Ee_min - me) + np.sqrt(me * (Ee_min + me) * (me * (Ee_min - me) + 2 * mmu ** 2)))


def sigma_muone(E, Ee_min):
    E = np.asarray(E, dtype=np.float64)
    s = me ** 2 + mmu ** 2 + 2 * me * E
    t_hi = 2.0 * me * (me - Ee_min)
    t_lo = -4.0 * ((s + me ** 2 - mmu ** 2) ** 2 / (4 * s) - me ** 2)
    pref = 16 * np.pi ** 2 * alpha_em ** 2 / (8.0 * np.pi * ((s - mmu ** 2) ** 2 + me ** 4 - 2 * (s + mmu ** 2) * me ** 2))
    a = -2.0 * (s ** 2 + mmu ** 4 + 5 * me ** 4 - 2 * me ** 2 * (2 * s + mmu ** 2)) * (1.0 / t_hi - 1.0 / t_lo)
    b = 2.0 * (s + 2 * mmu ** 2 - 2 * me ** 2) * np.log(t_hi / t_lo)
    return pref * (a + b + (t_hi - t_lo)) * _step(E - muone_threshold(Ee_min))
