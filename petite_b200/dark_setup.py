"""Host-side builder of the DarkShower constructor tables.

What the reference computes inside ``DarkShower.__init__`` with nested adaptive quadratures
(src/PETITE/dark_shower.py:254-593, using the cumulative interaction integrals of shower.py:298-354 and the
bound-electron cross-sections of atomic_annihilation.py / atomic_compton.py) and partly caches in
``dark_weights.pkl`` / ``dark_drate.pkl``.  Here ALL of it is written to ``dark_setup_<material>_mV<mV>.npz``:

    weights/<name>        (n,2)   integrated emission probability vs initial energy   (dark_shower.py:337-399)
    drate/<name>/E|table          10 energy bins of the emission rate per initial energy (dark_shower.py:454-493)
    nsdark/<P>/x|y                log10 nodes of n*sigma_dark(E)                      (dark_shower.py:294-309)

Only ``bound_electron=True`` is supported.  The ~6 000 adaptive quadratures run on the GPU (``pb_quad_batch``: QUADPACK's QAGS restated in
csrc/quadpack.cuh, one thread per integral, a fraction of a second); ``scipy_runner`` evaluates the same calls with
``scipy.integrate.quad`` as the reference does (30-60 s per (material, mV)) and is the CPU tests' yardstick.
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


# ------------------------------------------------------------------------------------------------ the quadratures
# Every integral of the set-up is one "call" (struct pb_quad_call, include/petite_b200.h) over a list of linear tables:
#   kind 0: f(E) = table[tab](E)                                       cumulative interaction integrals, shower.py:298-320
#   kind 1: f(E) = 10^table[tab](log10 E) / dEdx_cm * survival(E, Ei)   emission rate, dark_shower.py:311-335, shower.py:322-354
# A "runner" integrates a batch of calls.  The product runner is the GPU (pb_quad_batch: QUADPACK's QAGS restated in
# csrc/quadpack.cuh, one thread per integral); ``scipy_runner`` evaluates the same calls with scipy.integrate.quad on Python
# integrands (what the reference does) and is what the CPU tests pin the restatement to.
CALL_DTYPE = np.dtype([("kind", "<i4"), ("tab", "<i4"), ("surv", "<i4", (3,)), ("cut", "<i4"), ("Ei", "<f8"), ("a", "<f8"), ("b", "<f8")])


def _call(kind, tab, a, b, Ei=0.0, surv=(), cut=False):
    c = np.zeros((), dtype=CALL_DTYPE)
    c["kind"], c["tab"], c["cut"], c["Ei"], c["a"], c["b"] = kind, tab, int(cut), Ei, a, b
    c["surv"] = (list(surv) + [-1, -1, -1])[:3]
    return c


def integrand(tables, dEdx_m, c):
    """Python twin of QuadIntegrand (csrc/engine.cu) for one call; ``tables`` = list of LinearTable."""
    if c["kind"] == 0:
        t = tables[c["tab"]]
        return lambda E: t(E)
    ns, surv, Ei, cut = tables[c["tab"]], [tables[k] for k in c["surv"] if k >= 0], float(c["Ei"]), bool(c["cut"])
    dEdx_cm = dEdx_m * K.cmtom

    def f(E):
        v = 10 ** ns(np.log10(E))
        if cut and v < 1.0e-18:
            return 0.0
        d = sum(t(Ei) - t(E) for t in surv)
        if d < 0.0 or E > Ei:
            return 0.0
        return v / dEdx_cm * np.exp(-d / dEdx_m / K.cmtom)
    return f


def scipy_runner(tables, dEdx_m, calls):
    return np.array([quad(integrand(tables, dEdx_m, c), float(c["a"]), float(c["b"]), full_output=1)[0] for c in calls])


def c_abi_runner(fn, engine=None, on_error=None):
    """Runner over a C entry point with pb_quad_batch's table / call arguments (the engine's, or the host build used by the tests)."""
    import ctypes as C

    def run(tables, dEdx_m, calls):
        calls = np.ascontiguousarray(calls, dtype=CALL_DTYPE)
        n_t = len(tables)
        xs = [np.ascontiguousarray(t.x, dtype=np.float64) for t in tables]
        ys = [np.ascontiguousarray(t.y, dtype=np.float64) for t in tables]
        dp = C.POINTER(C.c_double)
        px = (dp * n_t)(*[a.ctypes.data_as(dp) for a in xs])
        py = (dp * n_t)(*[a.ctypes.data_as(dp) for a in ys])
        tn = np.array([len(a) for a in xs], dtype=np.int32)
        fill = np.array([float(t.fill_value) for t in tables], dtype=np.float64)
        out, err, ier = np.zeros(len(calls)), np.zeros(len(calls)), np.zeros(len(calls), dtype=np.int32)
        args = (n_t, tn.ctypes.data_as(C.POINTER(C.c_int32)), px, py, fill.ctypes.data_as(dp), C.c_double(dEdx_m),
                C.c_void_p(calls.ctypes.data), C.c_int64(len(calls)), out.ctypes.data_as(dp), err.ctypes.data_as(dp),
                ier.ctypes.data_as(C.POINTER(C.c_int32)))
        rc = fn(engine, *args) if engine is not None else fn(*args)
        if rc != 0:
            if on_error is not None:
                on_error(rc)
            raise RuntimeError(f"quadrature batch failed with code {rc}")
        run.last_ier = ier
        return out
    return run


def gpu_runner(sh):
    from . import _capi as capi
    return c_abi_runner(capi.lib.pb_quad_batch, sh._engine, on_error=lambda rc: capi.check(sh._engine, rc))


def build(sh, path, runner=None):
    """Compute the set-up tables for ``sh`` (a partly constructed DarkShower) and write them to ``path``.  ``runner`` integrates the
    batches of quadrature calls; default: the GPU engine of ``sh``."""
    runner = gpu_runner(sh) if runner is None else runner
    X, t = sh._xsec, sh._nsigma_tables
    nZ, ne = sh.get_n_targets()
    G = K.GeVsqcm2
    dEdx_m = sh._dEdx * 0.1                      # GeV/m

    # ---- phase 1 (shower.py:298-320): II(E_i) = int_{E_0}^{E_i} n*sigma dE on the table's own energy grid
    ii_spec = [("Brem", t["Brem"], X["Brem"][:, 0]), ("Ann", t["Ann"], X["Ann"][:, 0]), ("Moller", t["Moller"], t["Moller"].x),
               ("Bhabha", t["Bhabha"], t["Bhabha"].x),
               ("MuonBrem", t["MuonE"], X["MuonBrem"][:, 0]),          # SURVEY Q-4: integrates n*sigma_MuonE
               ("MuonE", t["MuonE"], X["MuonE"][:, 0])]
    tabs1 = [tb for _, tb, _ in ii_spec]
    calls1 = np.array([_call(0, k, float(grid[0]), float(e)) for k, (_, _, grid) in enumerate(ii_spec) for e in grid], dtype=CALL_DTYPE)
    y1 = runner(tabs1, dEdx_m, calls1)
    II, pos = {}, 0
    for name, _, grid in ii_spec:
        II[name] = LinearTable(np.asarray(grid, dtype=float), y1[pos:pos + len(grid)])
        pos += len(grid)

    DBS, DMB = sh._dark_brem_cross_section, sh._dark_muon_brem_cross_section
    E_res, E_thr = sh._resonant_annihilation_energy, sh._compton_threshold_energy
    mce = sh._minimum_calculable_energy
    Ea = np.logspace(np.log10(max(mce[-11], 0.001 * E_res)), np.log10(DBS[-1][0]), 200)            # dark_shower.py:254-264
    ann_bound = np.column_stack([Ea, [sigma_atomic_annihilation(e, sh._mV, sh.Zeff) for e in Ea]])
    Ec = np.logspace(np.log10(max(mce[22], 0.001 * E_thr)), np.log10(sh._dark_compton_cross_section[-1][0]), 200)
    comp_bound = np.column_stack([Ec, [sigma_atomic_compton(e, sh._mV, sh.Zeff) for e in Ec]])
    ns = {"DarkBrem": _LogLog(DBS[:, 0], nZ * G * DBS[:, 1]), "DarkAnn": _LogLog(ann_bound[:, 0], ne * G * ann_bound[:, 1]),
          "DarkComp": _LogLog(comp_bound[:, 0], ne * G * comp_bound[:, 1]), "DarkMuonBrem": _LogLog(DMB[:, 0], nZ * G * DMB[:, 1])}

    # ---- phase 2 (dark_shower.py:311-399, 454-493): emission weights and dRate/dE bins
    names2 = ["Brem", "Ann", "Moller", "Bhabha", "MuonBrem", "MuonE"]
    tabs2 = [II[n] for n in names2] + [ns[P]._t for P in ("DarkBrem", "DarkMuonBrem", "DarkAnn")]
    iid = {n: k for k, n in enumerate(names2)}
    nsid = {"DarkBrem": 6, "DarkMuonBrem": 7, "DarkAnn": 8}
    surv_e, surv_p, surv_mu = [iid["Brem"], iid["Moller"]], [iid["Brem"], iid["Ann"], iid["Bhabha"]], [iid["MuonBrem"], iid["MuonE"]]
    Eb, Em, Ean = DBS[:, 0], DMB[:, 0], ann_bound[:, 0]
    rates = {"brem_elec": (nsid["DarkBrem"], surv_e, True, Eb, 11), "brem_positron": (nsid["DarkBrem"], surv_p, True, Eb, -11),
             "muon_brem": (nsid["DarkMuonBrem"], surv_mu, True, Em, 13), "annihilation": (nsid["DarkAnn"], surv_p, False, Ean, -11)}
    calls2, layout = [], {}
    for name, (tab, surv, cut, Es, pid) in rates.items():
        first = len(calls2)
        edges_all = []
        for Ei in Es:                                                    # weights: two pieces, split 10 mean free paths below Ei
            Ei = float(Ei)
            reach = 10 * sh.get_mfp([pid, Ei]) * dEdx_m
            brk = Ei - reach
            brk = brk if brk > Es[0] else float(Es[0])
            calls2 += [_call(1, tab, float(Es[0]), brk, Ei, surv, cut), _call(1, tab, brk, Ei, Ei, surv, cut)]
        for Ei in Es:                                                    # dRate: ten equal bins over the last 10 mean free paths
            Ei = float(Ei)
            edges = np.linspace(max(Ei - 10 * sh.get_mfp([pid, Ei]) * dEdx_m, float(Es[0])), Ei, 11)
            edges_all.append(edges)
            calls2 += [_call(1, tab, float(edges[i]), float(edges[i + 1]), Ei, surv, cut) for i in range(10)]
        layout[name] = (first, len(Es), edges_all)
    y2 = runner(tabs2, dEdx_m, np.array(calls2, dtype=CALL_DTYPE))

    out = {"meta": np.array([sh._mV, sh._mV_estimator, E_res, E_thr, sh.g_e, sh.kinetic_mixing, sh.Zeff, 1.0])}
    for name, (tab, surv, cut, Es, pid) in rates.items():
        first, n, edges_all = layout[name]
        w = y2[first:first + 2 * n].reshape(n, 2)
        out[f"weights/{name}"] = np.column_stack([Es, w[:, 0] + w[:, 1]])
        d = y2[first + 2 * n:first + 12 * n].reshape(n, 10)
        centres = np.array([[(e[i] + e[i + 1]) / 2.0 for i in range(10)] for e in edges_all])
        out[f"drate/{name}/E"], out[f"drate/{name}/table"] = np.asarray(Es, dtype=float), np.stack([centres, d], axis=2)
    for P, tab in ns.items():
        out[f"nsdark/{P}/x"], out[f"nsdark/{P}/y"] = tab.lx, tab.ly
    md = sh._minimum_calculable_dark_energy
    pids, procs, Es = zip(*[(pid, pr, float(e)) for pid, d in md.items() for pr, e in d.items()])
    out["min_dark_pid"], out["min_dark_proc"], out["min_dark_E"] = np.array(pids), np.array(procs), np.array(Es)
    np.savez_compressed(path, **out)
    return path
