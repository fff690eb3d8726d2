"""Differential cross-sections in VEGAS-map variables (oracle).  TEST INFRASTRUCTURE.

Each function takes ``x`` with the sampled variables on the LAST axis (any leading
batch shape, or a single point) and an ``ev`` dict (``E_inc, Z_T, A_T, mT, mV,
Eg_min, Ee_min, m_lepton``) and returns dsigma (GeV^-2 per unit map volume).
They restate reference ``src/PETITE/all_processes.py`` (line ranges in each
docstring) and ``src/PETITE/radiative_return.py:26-79``; quirks Q-5, Q-19, Q-22 of
SURVEY.md are reproduced, not fixed.
"""
import numpy as np

from .consts import alpha_em, m_electron, m_muon, m_proton, GeV

PI = np.pi
_me = m_electron
_me2 = m_electron ** 2


def _cols(x, n):
    x = np.asarray(x, dtype=np.float64)
    return [x[..., i] for i in range(n)]


def form_factor_elastic(Z, t):
    """all_processes.py:92-103 (aa, g2_elastic)."""
    a0 = 184.15 * (2.718) ** -0.5 * Z ** (-1.0 / 3.0) / _me
    return Z ** 2 * a0 ** 4 * t ** 2 / (1 + a0 ** 2 * t) ** 2


def form_factor_el_inel_over_t2(Z, A, t):
    """all_processes.py:111-133 (Gelastic_inelastic_over_tsquared)."""
    mu_p = 2.79
    c1 = (111 * Z ** (-1 / 3) / _me) ** 2
    c2 = 0.164 * GeV ** 2 * A ** (-2 / 3)
    Gel = (1.0 / (1.0 + c1 * t)) ** 2 * (1 + t / c2) ** (-2)
    ap2 = (773.0 * Z ** (-2.0 / 3) / _me) ** 2
    Ginel = (Z / (c1 ** 2 * Z ** 2) * np.power(ap2 / (1.0 + ap2 * t), 2.0)
             * ((1.0 + (mu_p ** 2 - 1.0) * t / (4.0 * m_proton ** 2)) / (1.0 + t / 0.71) ** 4))
    return Z ** 2 * c1 ** 2 * (Gel + Ginel)


def ds_brem(x, ev):
    """all_processes.py:140-206 (dsigma_brem_dimensionless); also used for MuonBrem (Q-5)."""
    x1, x2, x3, x4 = _cols(x, 4)
    ep, Egmin, ml = ev["E_inc"], ev["Eg_min"], ev["m_lepton"]
    span = ep - ml - Egmin
    w = Egmin + x1 * span
    d = ep / (2 * ml) * (x2 + x3)
    dp = ep / (2 * ml) * (x2 - x3)
    ph = (x4 - 1 / 2) * 2 * PI
    epp = ep - w
    ok = (Egmin < w) & (w < ep - ml) & (ml < epp) & (epp < ep) & (d > 0.0) & (dp > 0.0)
    cph = np.cos(ph)
    with np.errstate(all="ignore"):
        qsq = ml ** 2 * ((d ** 2 + dp ** 2 - 2 * d * dp * cph)
                         + ml ** 2 * ((1 + d ** 2) / (2 * ep) - (1 + dp ** 2) / (2 * epp)) ** 2)
        PF = 8.0 / PI * alpha_em * (alpha_em / ml) ** 2 * (epp * ml ** 4) / (w * ep * qsq ** 2) * d * dp
        jac = PI * ep ** 2 * span / ml ** 2
        FF = form_factor_elastic(ev["Z_T"], qsq)
        T1 = d ** 2 / (1 + d ** 2) ** 2
        T2 = dp ** 2 / (1 + dp ** 2) ** 2
        T3 = w ** 2 / (2 * ep * epp) * (d ** 2 + dp ** 2) / ((1 + d ** 2) * (1 + dp ** 2))
        T4 = -(epp / ep + ep / epp) * (d * dp * cph) / ((1 + d ** 2) * (1 + dp ** 2))
        # Q-22: boolean mask is MULTIPLIED in (nan*0 = nan survives)
        return ok * PF * (T1 + T2 + T3 + T4) * jac * FF


def ds_pairprod(x, ev):
    """all_processes.py:532-622 (dsigma_pairprod_dimensionless)."""
    x1, x2, x3, x4 = _cols(x, 4)
    w = ev["E_inc"]
    epp = _me + x1 * (w - 2 * _me)
    dp = w / (2 * _me) * (x2 + x3)
    dm = w / (2 * _me) * (x2 - x3)
    ph = x4 * 2 * PI
    epm = w - epp
    ok = (_me < epm) & (epm < w) & (_me < epp) & (epp < w) & (dm > 0.0) & (dp > 0.0)
    cph = np.cos(ph)
    with np.errstate(all="ignore"):
        q2r = (dp ** 2 + dm ** 2 + 2.0 * dp * dm * cph) + _me2 * (
            (1.0 + dp ** 2) / (2.0 * epp) + (1.0 + dm ** 2) / (2.0 * epm)) ** 2
        PF = 8.0 / PI * alpha_em * (alpha_em / _me) ** 2 * epp * epm / (w ** 3 * q2r ** 2) * dp * dm
        jac = PI * w ** 2 * (w - 2 * _me) / _me2
        FF = form_factor_elastic(ev["Z_T"], _me2 * q2r)
        T1 = -1.0 * dp ** 2 / (1.0 + dp ** 2) ** 2
        T2 = -1.0 * dm ** 2 / (1.0 + dm ** 2) ** 2
        T3 = w ** 2 / (2.0 * epp * epm) * (dp ** 2 + dm ** 2) / ((1.0 + dp ** 2) * (1.0 + dm ** 2))
        T4 = (epp / epm + epm / epp) * (dp * dm * cph) / ((1.0 + dp ** 2) * (1.0 + dm ** 2))
        return np.where(ok, PF * (T1 + T2 + T3 + T4) * jac * FF, 0.0)


def ds_compton(x, ev):
    """all_processes.py:625-742 (dsigma_compton_dCT); mV>0 is DarkComp."""
    (ct,) = _cols(x, 1)
    Eg, mV = ev["E_inc"], ev.get("mV", 0.0)
    s = _me2 + 2 * Eg * _me
    if s < (_me + mV) ** 2:
        return np.zeros_like(ct)
    lam = np.sqrt((s - mV ** 2) ** 2 - 2 * _me2 * (s + mV ** 2) + _me ** 4)
    jac = (s - _me2) / (2 * s) * lam
    lam2 = np.sqrt(_me ** 4 + (mV ** 2 - s) ** 2 - 2 * _me2 * (mV ** 2 + s))
    t = -1 / 2 * (_me ** 4 + s * (-(mV ** 2) + s + ct * lam2) - _me2 * (mV ** 2 + 2 * s + ct * lam2)) / s
    PF = 2.0 * PI * alpha_em ** 2 / (s - _me2) ** 2
    if mV == 0.0:
        T1 = (6.0 * _me2 * s + 3.0 * _me ** 4 - s ** 2) / ((_me2 - s) * (-_me2 + s + t))
        T2 = 4 * _me ** 4 / (s + t - _me2) ** 2
        T3 = (t * (s - _me2) + (s + _me2) ** 2) / (s - _me2) ** 2
    else:
        T1 = (2.0 * _me2 * (mV ** 2 - 3 * s) - 3 * _me ** 4 - 2 * mV ** 2 * s + 2 * mV ** 4 + s ** 2) / (
            (_me2 - s) * (_me2 + mV ** 2 - s - t))
        T2 = (2 * _me2 * (2 * _me2 + mV ** 2)) / (_me2 + mV ** 2 - s - t) ** 2
        T3 = ((_me2 + s) * (_me2 + mV ** 2 + s) + t * (s - _me2)) / (_me2 - s) ** 2
    return PF * jac * (T1 + T2 + T3)


def ds_annihilation(x, ev):
    """all_processes.py:469-529 (dsigma_annihilation_dCT)."""
    (ct,) = _cols(x, 1)
    Ee, mV, EgMin = ev["E_inc"], ev.get("mV", 0.0), ev.get("Eg_min", 0.0)
    s = 2.0 * _me * (Ee + _me)
    with np.errstate(all="ignore"):
        ctMax = (np.sqrt((Ee + _me) / (Ee - _me)) * (2 * _me * (Ee - 2 * EgMin + _me) - mV ** 2)
                 / (2 * _me * (Ee + _me) - mV ** 2))
    if s < mV ** 2:
        return np.zeros_like(ct)
    b = np.sqrt(1.0 - 4.0 * _me2 / s)
    val = (4.0 * PI * alpha_em ** 2 / (s * (1 - b ** 2 * ct ** 2))
           * ((s - mV ** 2) / (2 * s) * (1 + ct ** 2) + 2.0 * mV ** 2 / (s - mV ** 2)))
    return np.where(ct > ctMax, 0.0, val)


def ds_moller(x, ev):
    """all_processes.py:787-840 (dsigma_moller_dCT)."""
    (ct,) = _cols(x, 1)
    Ee, DE = ev["E_inc"], ev.get("Ee_min", 0.010)
    lim = 2.0 * DE / (Ee - _me)
    ok = (ct > -1 + lim) & (ct < 1.0 - lim)
    s = _me2 + 2 * Ee * _me
    with np.errstate(all="ignore"):
        val = (16 * PI ** 2 * alpha_em ** 2
               * (s ** 2 * (3 + ct ** 2) ** 2 - 8 * _me2 * s * (7 + ct ** 4)
                  + 16 * _me ** 4 * (6 - 3 * ct ** 2 + ct ** 4))
               / (8 * PI * s * (s - 4 * _me2) ** 2 * (1 - ct) ** 2 * (1 + ct) ** 2))
    return np.where(ok, val, 0.0)


def ds_bhabha(x, ev):
    """all_processes.py:948-1007 (dsigma_bhabha_dCT)."""
    (ct,) = _cols(x, 1)
    Ee, DE = ev["E_inc"], ev.get("Ee_min", 0.010)
    lim = 2.0 * DE / (Ee - _me)
    ok = (ct > -1 + lim) & (ct < 1.0 - lim)
    s = _me2 + 2 * Ee * _me
    m = _me
    with np.errstate(all="ignore"):
        num = (256 * (-1 + ct) ** 2 * ct ** 2 * m ** 8
               - 128 * (-1 + ct) * (1 + ct * (1 + ct) * (-3 + 2 * ct)) * m ** 6 * s
               + 16 * (7 + ct * (2 + ct * (-5 + 6 * (-1 + ct) * ct))) * m ** 4 * s ** 2
               - 8 * (7 + ct * (-3 + ct * (3 + ct * (-1 + 2 * ct)))) * m ** 2 * s ** 3
               + (3 + ct ** 2) ** 2 * s ** 4)
        val = (alpha_em ** 2 * PI * num) / (2 * (-1 + ct) ** 2 * s ** 3 * (-4 * m ** 2 + s) ** 2)
    return np.where(ok, val, 0.0)


def ds_muone(x, ev):
    """all_processes.py:843-886 (dsigma_muonelectron_dCT)."""
    (ct,) = _cols(x, 1)
    Emu, DE = ev["E_inc"], ev.get("Ee_min", 0.010)
    s = _me2 + m_muon ** 2 + 2 * _me * Emu
    t_limit = 2.0 * _me * (_me - DE)
    t = -2.0 * (1 - ct) * ((s + _me2 - m_muon ** 2) ** 2 / (4.0 * s) - _me2)
    with np.errstate(all="ignore"):
        val = (16 * PI ** 2 * alpha_em ** 2
               * (s ** 2 + 2 * (_me2 + m_muon ** 2) * (2 * t + m_muon ** 2 - 3 * _me2)
                  + (s + t - 4 * _me2) ** 2) / (16 * PI * s * t ** 2))
    return np.where(t < t_limit, val, 0.0)


def ds_darkbrem(x, ev):
    """all_processes.py:208-372 (dsig_dx_dcostheta_dark_brem_exact_tree_level, Method='Log')."""
    xx, l1mct, lttilde = _cols(x, 3)
    ml, mV, Eb, MT = ev["m_lepton"], ev["mV"], ev["E_inc"], ev["mT"]
    with np.errstate(all="ignore"):
        omc = 10 ** l1mct
        cth = 1.0 - omc
        ttilde = 10 ** lttilde
        Jac = omc * ttilde * np.log(10.0) ** 2
        k = np.sqrt(np.fabs((xx * Eb) ** 2 - mV ** 2))
        p = np.sqrt(Eb ** 2 - ml ** 2)
        V = np.sqrt(p ** 2 + k ** 2 - 2 * p * k * cth)
        utilde = -2 * (xx * Eb ** 2 - k * p * cth) + mV ** 2
        Er = (1 - xx) * Eb + MT
        discr = utilde ** 2 + 4 * MT * utilde * Er + 4 * MT ** 2 * V ** 2
        sq = np.sqrt(np.abs(discr))
        den = 2 * Er ** 2 - 2 * V ** 2
        Qp = np.fabs((V * (utilde + 2 * MT * Er) + Er * sq) / den)
        Qm = np.fabs((V * (utilde + 2 * MT * Er) - Er * sq) / den)
        tplus = 2 * MT * (np.sqrt(MT ** 2 + Qp ** 2) - MT)
        tminus = 2 * MT * (np.sqrt(MT ** 2 + Qm ** 2) - MT)
        tconv = (2 * MT * (MT + Eb) * np.sqrt(Eb ** 2 + ml ** 2) / (MT * (MT + 2 * Eb) + ml ** 2)) ** 2
        t = ttilde * tconv
        q0 = -t / (2 * MT)
        q = np.sqrt(t ** 2 / (4 * MT ** 2) + t)
        cthq = -(V ** 2 + q ** 2 + ml ** 2 - (Eb + q0 - xx * Eb) ** 2) / (2 * V * q)
        mm = mV ** 2 + 2 * ml ** 2
        Am2 = -8 * MT * (4 * Eb ** 2 * MT - t * (2 * Eb + MT)) * mm
        A1 = 8 * MT ** 2 / utilde
        Am1 = (8 / utilde) * (
            MT ** 2 * (2 * t * utilde + utilde ** 2
                       + 4 * Eb ** 2 * (2 * (xx - 1) * mm - t * ((xx - 2) * xx + 2))
                       + 2 * t * (-(mV ** 2) + 2 * ml ** 2 + t))
            - 2 * Eb * MT * t * ((1 - xx) * utilde + (xx - 2) * (mm + t))
            + t ** 2 * (utilde - mV ** 2))
        A0 = (8 / utilde ** 2) * (
            MT ** 2 * (2 * t * utilde + (t - 4 * Eb ** 2 * (xx - 1) ** 2) * mm)
            + 2 * Eb * MT * t * (utilde - (xx - 1) * mm))
        Y = -t + 2 * q0 * Eb - 2 * q * p * (p - k * cth) * cthq / V
        W = np.fabs(Y ** 2 - 4 * q ** 2 * p ** 2 * k ** 2 * (1 - cth ** 2) * (1 - cthq ** 2) / V ** 2)
        ok = ((xx * Eb >= mV) & (discr >= 0) & (tplus > tminus) & (t > tminus) & (t < tplus)
              & (np.fabs(cthq) <= 1.0) & (W > 0))
        phi_int = np.where(ok, (A0 + Y * A1 + Am1 / np.sqrt(W) + Y * Am2 / W ** 1.5) / (8 * MT ** 2), 0.0)
        FF = form_factor_el_inel_over_t2(ev["Z_T"], ev["A_T"], t)
        ans = FF * np.power(alpha_em, 3) * k * Eb * phi_int / (p * np.sqrt(k ** 2 + p ** 2 - 2 * p * k * cth))
        return np.where(ok, ans * tconv * Jac, 0.0)


def kf_beta(s):
    return (2.0 * alpha_em / PI) * (np.log(s / _me2) - 1.0)


def fl_kf(x, s):
    """radiative_return.py:26-35."""
    beta = kf_beta(s)
    x = np.where(x >= 1.0, 1.0 - 1e-10, x)
    return (beta / 16.0) * ((8.0 + 3.0 * beta) * np.power(1.0 - x, beta / 2.0 - 1.0) - 4.0 * (1.0 + x))


def fl_kf_scaled(x, s):
    """radiative_return.py:37-45."""
    beta = kf_beta(s)
    return (beta / 16.0) * ((8.0 + 3.0 * beta) - 4.0 * (1.0 + x) * np.power(1.0 - x, 1.0 - beta / 2.0))


def transformed_lumi_integrand(s, y, u):
    """radiative_return.py:61-79."""
    beta = kf_beta(s)
    x = 1.0 - np.power(u, 2.0 / beta)
    return fl_kf(y / x, s) * fl_kf_scaled(x, s) * (1.0 / x) * (2.0 / beta)


def ds_darkann(x, ev):
    """all_processes.py:400-466 (dsigma_radiative_return_du)."""
    (u0,) = _cols(x, 1)
    mV, Ee = ev["mV"], ev["E_inc"]
    s = 2.0 * _me * (Ee + _me)
    if s < mV ** 2:
        return np.zeros_like(u0)
    beta = kf_beta(s)
    umax = np.power(1.0 - mV ** 2 / s, beta / 2.0)
    betaf = np.sqrt(1.0 - 4.0 * _me2 / mV ** 2)
    prefac = (4.0 * PI ** 2) * alpha_em * betaf * (3.0 / 2.0 - betaf ** 2 / 2.0) / s * umax
    with np.errstate(all="ignore"):
        x1 = 1.0 - np.power(u0 * umax, 2.0 / beta)
        x2 = mV ** 2 / (x1 * s)
        ok = (x2 < 1.0) & (x1 > 0.0) & (u0 < 1.0)
        val = 2.0 * prefac * transformed_lumi_integrand(s, mV ** 2 / s, u0 * umax)
    return np.where(ok, val, 0.0)


DSIGMA = {"Brem": ds_brem, "MuonBrem": ds_brem, "PairProd": ds_pairprod, "Comp": ds_compton,
          "Ann": ds_annihilation, "Moller": ds_moller, "Bhabha": ds_bhabha, "MuonE": ds_muone,
          "DarkBrem": ds_darkbrem, "DarkMuonBrem": ds_darkbrem, "DarkAnn": ds_darkann,
          "DarkComp": ds_compton}


# ---- analytic totals used to build n*sigma tables (setup only) ----

def sigma_moller(E, Ee_min):
    """all_processes.py:907-944."""
    E = np.asarray(E, dtype=np.float64)
    T = Ee_min - _me
    thr = 3 * _me + 4 * T
    on = np.heaviside(E - thr, 1)
    off = np.heaviside(thr - E, 1)
    PF = 2 * PI * alpha_em ** 2 / (_me * (E ** 2 - _me2))
    T1 = (E - 3 * _me - 4 * T + 2 * E ** 2 * (-2 / (E - 3 * _me - 2 * T) + 1 / T
                                              + 1 / (-E + _me + T) + 2 / (E + _me + 2 * T)))
    with np.errstate(all="ignore"):
        T2 = (2 * _me * (_me - 2 * E) / (E - _me)
              * np.log(((-E + _me + T) * (-E + 3 * _me + 2 * T) / (T * (E + _me + 2 * T))) * on + off))
    return PF * (T1 + T2) * on


def sigma_bhabha(E, Ee_min):
    """all_processes.py:1010-1051."""
    E = np.asarray(E, dtype=np.float64)
    m = _me
    T = Ee_min - m
    thr = 3 * m + 4 * T
    on = np.heaviside(E - thr, 1)
    off = np.heaviside(thr - E, 1)
    PF = PI * alpha_em ** 2 / (12 * (E - m) * m * (E + m) ** 3 * (E - 3 * m - 2 * T) * T)
    T1 = (E - 3 * m - 4 * T) * (24 * E ** 2 * (E + m) ** 2
                                + (E - 3 * m) * (31 * E ** 2 + 84 * E * m + 57 * m ** 2) * T
                                - 4 * (16 * E ** 2 + 39 * E * m + 33 * m ** 2) * T ** 2
                                + 8 * (E - 3 * m) * T ** 3 - 8 * T ** 4)
    with np.errstate(all="ignore"):
        T2 = (24 * (E + m) * (2 * E ** 2 + 4 * E * m + m ** 2) * (E - 3 * m - 2 * T) * T
              * np.log((2 * T / (E - 3 * m - 2 * T)) * on + off))
    return PF * (T1 + T2) * on


def muone_threshold(Ee_min):
    """all_processes.py:893 / shower.py:289."""
    return 1.0 / (2.0 * _me) * (_me * (Ee_min - _me)
                                + np.sqrt(_me * (Ee_min + _me) * (_me * (Ee_min - _me) + 2 * m_muon ** 2)))


def sigma_muone(E, Ee_min):
    """all_processes.py:888-905."""
    E = np.asarray(E, dtype=np.float64)
    thr = muone_threshold(Ee_min)
    s = _me2 + m_muon ** 2 + 2 * _me * E
    t_max = 2.0 * _me * (_me - Ee_min)
    t_min = -4.0 * ((s + _me2 - m_muon ** 2) ** 2 / (4 * s) - _me2)
    PF = 16 * PI ** 2 * alpha_em ** 2 / (8.0 * PI * ((s - m_muon ** 2) ** 2 + _me ** 4 - 2 * (s + m_muon ** 2) * _me2))
    T1 = -2.0 * (s ** 2 + m_muon ** 4 + 5 * _me ** 4 - 2 * _me2 * (2 * s + m_muon ** 2)) * (1.0 / t_max - 1.0 / t_min)
    T2 = 2.0 * (s + 2 * m_muon ** 2 - 2 * _me2) * np.log(t_max / t_min)
    T3 = t_max - t_min
    return PF * (T1 + T2 + T3) * np.heaviside(E - thr, 1)
