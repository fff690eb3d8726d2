"""Host-side builder of the DarkShower constructor tables.

What the reference computes inside ``DarkShower.__init__`` with nested adaptive quadratures
(src/PETITE/dark_shower.py:254-593, using the cumulative interaction integrals of shower.py:298-354 and the
bound-electron cross-sections of atomic_annihilation.py / atomic_compton.py) and partly caches in
``dark_weights.pkl`` / ``dark_drate.pkl``.  Here ALL of it is written to ``dark_setup_<material>_mV<mV>.npz``:

    weights/<name>        (n,2)   integrated emission probability vs initial energy   (dark_shower.py:337-399)
    drate/<name>/E|table          10 energy bins of the emission rate per initial energy (dark_shower.py:454-493)
    nsdark/<P>/x|y                log10 nodes of n*sigma_dark(E)                      (dark_shower.py:294-309)

Only ``bound_electron=True`` is supported.  It is set-up code: scalar SciPy quadrature, ~30-60 s per (material, mV).
"""
import numpy as np
from scipy.integrate import quad

from . import constants as K
from .shower import LinearTable

# atomic_annihilation.py:4-5 keeps its own rounded constants (SURVEY Q-16); atomic_compton.py uses the global ones
_AA_ALPHA, _AA_ME = 1 / 137, 0.511e-3


def _poly(c, x):
    """sum_k c[k] x^k, evaluated term by term (matches the reference's explicit sums to rounding)."""
    return sum(ck * x ** k for k, ck in enumerate(c))


def _isr_tail_integral(a, b):
    """Closed-form b-integral of the radiative tail (atomic_annihilation.py:10-68, "fancy_integral")."""
    A = 1 + a ** 2
    if b >= 1000:      # large-b series (atomic_annihilation.py:64-68)
        return (3081 / 280 - (3081 * a ** 2) / 40) / b ** 8 - (709 * a) / (35 * b ** 7) - 91 / (30 * b ** 6)
    t1 = A * _poly([a * A ** 3 * (5 + 3 * a ** 2), -(A ** 2) * (-1 + 15 * a ** 2 + 12 * a ** 4),
                    a * A * (-23 + 38 * a ** 2 + 21 * a ** 4), 4 * (5 + a ** 2 - 18 * a ** 4 - 6 * a ** 6),
                    3 * a * (3 + a ** 2) * (-1 + 7 * a ** 2), 7 - 29 * a ** 2 - 12 * a ** 4, a * (7 + 3 * a ** 2)], b)
    t2 = _poly([3 * A ** 6, -a * A ** 3 * (3 + 26 * a ** 2 + 15 * a ** 4), 3 * A ** 2 * (-7 + 21 * a ** 2 + 23 * a ** 4 + 11 * a ** 6),
                -3 * a * A * (7 + 69 * a ** 2 + 45 * a ** 4 + 15 * a ** 6), -11 + 132 * a ** 2 + 9 * a ** 4 * (38 + 5 * a ** 2 * (4 + a ** 2)),
                -3 * a * (11 + 73 * a ** 2 + 41 * a ** 4 + 11 * a ** 6), 3 * (-1 + 27 * a ** 2 + 17 * a ** 4 + 5 * a ** 6),
                -a * (15 + 10 * a ** 2 + 3 * a ** 4)], b)
    t4 = _poly([A ** 3, 6 * a * A ** 2, -3 * (1 + 6 * a ** 2 + 5 * a ** 4), 4 * a * (3 + 5 * a ** 2), -3 * (1 + 5 * a ** 2), 6 * a, -1], b)
    num = (t1 + t2 * np.arctan2(a, 1) + t2 * np.arctan2(1 + a ** 2 - a * b, b)
           + 4 * b * (t4 * (np.log(1 + (a - b) ** 2) - 2 * np.log(b))))
    return num / (8 * A ** 3 * (1 + (a - b) ** 2) ** 3 * b)


def sigma_atomic_annihilation(k, mV, Zeff):
    """e+ e-(bound) -> V: tree level + radiative tail (atomic_annihilation.py:70-105)."""
    al, me = _AA_ALPHA, _AA_ME
    lam = Zeff * al * me
    a, b = me / lam, mV ** 2 / (2 * k * lam)
    pref = (4.0 * np.pi * al) * ((mV ** 2 + 2 * me ** 2) * (2 / 3 / lam / me) * (1 / k ** 2))
    beta = 2 * al / np.pi * (np.log((2 * k * me + me ** 2) / me ** 2) - 1.0)
    return pref * (1 / ((a - b) ** 2 + 1) ** 3) + pref * (beta / 2) * _isr_tail_integral(a, b)


def _compton_bound_shape(a, b):
    """atomic_compton.py:6-92 ("combine"): two analytic branches plus a large-b series."""
    delta = (a - b) ** 2 + 1
    c3 = 3 * (a ** 2 + 1) ** 4 - 2 * a * (3 * a ** 4 + 10 * a ** 2 + 15) * (a ** 2 + 1) * b + 6 * (a ** 6 + 5 * a ** 4 + 15 * a ** 2 - 5) * b ** 2
    inner = (-a * (a ** 2 + 1) ** 3 * (3 * a ** 2 + 5) + 6 * (a ** 4 + 4 * a ** 2 - 5) * b ** 5 - 4 * a * (6 * a ** 4 + 23 * a ** 2 - 31) * b ** 4
             + (a - 1) * (a + 1) * (39 * a ** 4 + 190 * a ** 2 + 55) * b ** 3 - a * (a ** 2 + 1) * (33 * a ** 4 + 98 * a ** 2 - 111) * b ** 2
             + (a ** 2 + 1) ** 2 * (15 * a ** 4 + 32 * a ** 2 - 23) * b)
    logt = np.log(delta) - 2 * np.log(b)
    if a ** 2 + 1 < a * b:
        num = 8 * b * (a ** 2 - 6 * a * b + 1) * logt + (a ** 2 + 1) * inner / delta ** 2 + c3 * (np.arctan(b / (a ** 2 - a * b + 1)) + np.arctan(1 / a))
        e1 = -num / (8 * (a ** 2 + 1) ** 4 * b)
        if e1 < 0.0 or (a ** 2 + 1 < 0.01 * a * b):
            e1 = (-145 + 1015 * a ** 2 + 330 * a * b + 64 * b ** 2) / (420 * b ** 8)
        return e1
    term4 = delta ** 2 * (8 * b * (a ** 2 - 6 * a * b + 1) * logt + c3 * np.arctan(b / (a ** 2 - a * b + 1)))
    num = (a ** 2 + 1) * (-inner) + np.pi * c3 * delta ** 2 - c3 * delta ** 2 * np.arctan(1 / a) - term4
    return num / (8 * (a ** 2 + 1) ** 4 * b * delta ** 2)


def sigma_atomic_compton(k, mV, Zeff):
    """gamma e-(bound) -> V e- (atomic_compton.py:94-110)."""
    al, me = K.alpha_em, K.m_electron
    lam = Zeff * al * me
    a, b = me / lam, mV ** 2 / (2 * k * lam)
    beta = max(2 * al / np.pi * (np.log((2 * k * me + me ** 2) / me ** 2) - 1.0), 0.0)
    return (4.0 * np.pi * al) * ((mV ** 2 + 2 * me ** 2) * (2 / 3 / lam / me) * (1 / k ** 2) * (beta / 4) * _compton_bound_shape(a, b))


def _log_table(x, y):
    return np.log10(np.asarray(x) + 1e-20), np.log10(np.asarray(y) + 1e-20)


class _LogLog:
    def __init__(self, x, y):
        self.lx, self.ly = _log_table(x, y)
        self._t = LinearTable(self.lx, self.ly, fill_value=-20.0)

    def __call__(self, E):
        return 10 ** self._t(np.log10(E))


def _cumulative(table, grid):
    """shower.py:298-320: II(E_i) = int_{E_0}^{E_i} n*sigma dE on the table's own energy grid."""
    y = np.array([quad(table, grid[0], e, full_output=1)[0] for e in grid])
    return LinearTable(grid, y)


def build(sh, path):
    """Compute the set-up tables for ``sh`` (a partly constructed DarkShower) and write them to ``path``."""
    X, t = sh._xsec, sh._nsigma_tables
    nZ, ne = sh.get_n_targets()
    G = K.GeVsqcm2
    dEdx_m = sh._dEdx * 0.1                      # GeV/m
    dEdx_cm = dEdx_m * K.cmtom                   # GeV/cm

    II = {P: _cumulative(t[P], X[P][:, 0]) for P in ("Brem", "Ann")}
    II["Moller"] = _cumulative(t["Moller"], t["Moller"].x)
    II["Bhabha"] = _cumulative(t["Bhabha"], t["Bhabha"].x)
    II["MuonBrem"] = _cumulative(t["MuonE"], X["MuonBrem"][:, 0])     # SURVEY Q-4: integrates n*sigma_MuonE
    II["MuonE"] = _cumulative(t["MuonE"], X["MuonE"][:, 0])

    def survive(names):
        def f(E, Ei):                            # shower.py:322-354
            d = sum(II[n](Ei) - II[n](E) for n in names)
            if d < 0.0 or E > Ei:
                return 0.0
            return np.exp(-d / dEdx_m / K.cmtom)
        return f
    surv_e, surv_p, surv_mu = survive(("Brem", "Moller")), survive(("Brem", "Ann", "Bhabha")), survive(("MuonBrem", "MuonE"))

    DBS, DMB = sh._dark_brem_cross_section, sh._dark_muon_brem_cross_section
    E_res, E_thr = sh._resonant_annihilation_energy, sh._compton_threshold_energy
    mce = sh._minimum_calculable_energy
    Ea = np.logspace(np.log10(max(mce[-11], 0.001 * E_res)), np.log10(DBS[-1][0]), 200)            # dark_shower.py:254-264
    ann_bound = np.column_stack([Ea, [sigma_atomic_annihilation(e, sh._mV, sh.Zeff) for e in Ea]])
    Ec = np.logspace(np.log10(max(mce[22], 0.001 * E_thr)), np.log10(sh._dark_compton_cross_section[-1][0]), 200)
    comp_bound = np.column_stack([Ec, [sigma_atomic_compton(e, sh._mV, sh.Zeff) for e in Ec]])

    ns = {"DarkBrem": _LogLog(DBS[:, 0], nZ * G * DBS[:, 1]), "DarkAnn": _LogLog(ann_bound[:, 0], ne * G * ann_bound[:, 1]),
          "DarkComp": _LogLog(comp_bound[:, 0], ne * G * comp_bound[:, 1]), "DarkMuonBrem": _LogLog(DMB[:, 0], nZ * G * DMB[:, 1])}

    def rate(nsig, surv, cut):
        def f(E, Ei):                            # dark_shower.py:311-335
            v = nsig(E)
            if cut and v < 1.0e-18:
                return 0.0
            return v / dEdx_cm * surv(E, Ei)
        return f
    f_be, f_bp = rate(ns["DarkBrem"], surv_e, True), rate(ns["DarkBrem"], surv_p, True)
    f_mu, f_an = rate(ns["DarkMuonBrem"], surv_mu, True), rate(ns["DarkAnn"], surv_p, False)

    def weights(f, Es, pid):                     # dark_shower.py:337-399
        out = []
        for Ei in Es:
            brk = Ei - 10 * sh.get_mfp([pid, Ei]) * dEdx_m
            brk = brk if brk > Es[0] else Es[0]
            out.append(quad(f, Es[0], brk, args=(Ei), full_output=1)[0] + quad(f, brk, Ei, args=(Ei), full_output=1)[0])
        return np.column_stack([Es, out])

    def drate(f, Es, pid, floor):                # dark_shower.py:454-493
        tabs = []
        for Ei in Es:
            edges = np.linspace(max(Ei - 10 * sh.get_mfp([pid, Ei]) * dEdx_m, floor), Ei, 11)
            centres = np.array([(edges[i] + edges[i + 1]) / 2.0 for i in range(10)])
            tabs.append(np.column_stack([centres, [quad(f, edges[i], edges[i + 1], args=(Ei), full_output=1)[0] for i in range(10)]]))
        return np.asarray(Es, dtype=float), np.stack(tabs)

    Eb, Em, Ean = DBS[:, 0], DMB[:, 0], ann_bound[:, 0]
    out = {"meta": np.array([sh._mV, sh._mV_estimator, E_res, E_thr, sh.g_e, sh.kinetic_mixing, sh.Zeff, 1.0]),
           "weights/brem_elec": weights(f_be, Eb, 11), "weights/brem_positron": weights(f_bp, Eb, -11),
           "weights/muon_brem": weights(f_mu, Em, 13), "weights/annihilation": weights(f_an, Ean, -11)}
    for name, (f, Es, pid, floor) in {"brem_elec": (f_be, Eb, 11, Eb[0]), "brem_positron": (f_bp, Eb, -11, Eb[0]),
                                      "muon_brem": (f_mu, Em, 13, Em[0]), "annihilation": (f_an, Ean, -11, Ean[0])}.items():
        out[f"drate/{name}/E"], out[f"drate/{name}/table"] = drate(f, Es, pid, floor)
    for P, tab in ns.items():
        out[f"nsdark/{P}/x"], out[f"nsdark/{P}/y"] = tab.lx, tab.ly
    md = sh._minimum_calculable_dark_energy
    pids, procs, Es = zip(*[(pid, pr, float(e)) for pid, d in md.items() for pr, e in d.items()])
    out["min_dark_pid"], out["min_dark_proc"], out["min_dark_E"] = np.array(pids), np.array(procs), np.array(Es)
    np.savez_compressed(path, **out)
    return path
