"""Deterministic pieces of the CUDA path vs the reference's golden vectors and the oracle (fp64, 1e-12 relative)."""
import numpy as np
import pytest

from tests.gpu_util import shower, probe

pytestmark = pytest.mark.gpu

SM = ["Brem", "Ann", "PairProd", "Comp", "Moller", "Bhabha", "MuonE", "MuonBrem"]
CODE = {p: i for i, p in enumerate(SM)}
DIM = {"Brem": 4, "PairProd": 4, "MuonBrem": 4}


def test_philox_matches_oracle():
    from petite_b200 import _capi as capi
    from oracle import philox as ph
    sh = shower()
    rng = np.random.default_rng(5)
    a = rng.integers(0, 2 ** 32, size=(512, 6)).astype(np.float64)
    a[:4] = [[0, 0, 0, 0, 0, 0], [0xffffffff] * 6, [1, 2, 3, 4, 5, 6], [0xa4093822, 0x299f31d0, 0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344]]
    got = probe(sh, capi.PROBE_PHILOX, 0, a, 2)
    ai = a.astype(np.uint64)
    w0, w1 = ph.draw2((ai[:, 0], ai[:, 1]), ai[:, 2], ai[:, 3], ai[:, 4], ai[:, 5])
    assert np.array_equal(got[:, 0], w0) and np.array_equal(got[:, 1], w1)     # bit-exact


def test_hot_math():
    """The constant-memory elementary functions of the sub-step loop (physics.cuh hot_*) against libm: <= 2 ulp on the
    ranges the loop produces (log of (0, 1] and of [1, inf); exp(-x) on [1/20, 1/6]; sincos on [-pi/4, pi/4] and the libdevice
    fall-back beyond; sincos(2 pi u)); the branch-free reciprocal, square root and reciprocal square root <= 1 ulp."""
    from petite_b200 import _capi as capi
    sh = shower()
    rng = np.random.default_rng(11)
    n = 40000
    xlog = np.concatenate([1.0 - rng.random(n // 4), 2.0 ** -rng.uniform(0, 52, n // 4), 1.0 + 10.0 ** rng.uniform(-9, 9, n // 4),
                           10.0 ** rng.uniform(-300, 300, n // 4)])
    xlog[:3] = [1.0, 2.0 ** -52, 0.5]
    xexp = rng.uniform(1 / 20, 1 / 6, n); xexp[:2] = [1 / 20, 1 / 6]
    th = np.concatenate([rng.uniform(-np.pi / 4, np.pi / 4, n // 2), rng.normal(0, 1e-4, n // 4), rng.uniform(-50, 50, n // 4)])
    u = rng.random(n); u[:5] = [0.0, 0.25, 0.5, 0.75, 1 - 2.0 ** -52]
    got = probe(sh, capi.PROBE_HOTMATH, 0, np.column_stack([xlog, xexp, th, u]), 12)
    ulps = lambda a, b: np.abs(a - b) / np.spacing(np.maximum(np.abs(b), 1e-300))
    assert got[0, 0] == 0.0                                                   # log(1) is exactly 0
    lg = np.log(xlog)
    assert np.max(ulps(got[:, 0], lg)[lg != 0]) <= 2
    assert np.max(ulps(got[:, 1], np.exp(-xexp))) <= 2
    assert np.max(ulps(got[:, 2], np.sin(th))) <= 2 and np.max(ulps(got[:, 3], np.cos(th))) <= 2
    # sin / cos (2 pi u): compared in absolute terms (numpy rounds the argument 2 pi u, the kernel reduces exactly)
    L = np.longdouble
    a = 2 * L(np.pi) * u.astype(L) + 2 * L(1.2246467991473532e-16) * u.astype(L)
    assert np.max(np.abs(got[:, 4] - np.sin(a).astype(np.float64))) < 4e-16
    assert np.max(np.abs(got[:, 5] - np.cos(a).astype(np.float64))) < 4e-16
    assert np.max(ulps(got[:, 6], 1.0 / xlog)) <= 1
    # branch-free square root / reciprocal square root (MUFU.RSQ64H + Newton): <= 1 ulp; fast_sqrt0(0) is exactly 0
    xl = xlog.astype(L)                           # 80-bit yardstick: 1 / np.sqrt(x) is itself rounded twice
    assert np.max(np.abs(got[:, 7].astype(L) - np.sqrt(xl)) / np.spacing(np.sqrt(xlog))) <= 1
    assert np.max(np.abs(got[:, 8].astype(L) - 1 / np.sqrt(xl)) / np.spacing(1.0 / np.sqrt(xlog))) <= 1.01
    assert np.all(got[:, 9] == 0.0)
    # cos(pi t) of the sampler integrands (hot_cospi), t = 2 u - 1 (Brem) and 2 u (PairProd): <= 2 ulp where |cos| is not small, and in
    # absolute terms everywhere (the reference rounds (u - 1/2) 2 pi first, the kernel reduces exactly)
    for col, t in ((10, 2.0 * u - 1.0), (11, 2.0 * u)):
        tl = t.astype(L)
        exact = np.cos(L(np.pi) * tl + L(1.2246467991473532e-16) * tl)
        err = np.abs(got[:, col].astype(L) - exact)
        assert np.max(err) < 2.3e-16
        big = np.abs(exact) > 0.5
        assert np.max(err[big] / np.spacing(np.abs(exact[big]).astype(np.float64))) <= 2
    assert got[0, 10] == -1.0 and got[0, 11] == 1.0 and got[2, 10] == 1.0 and got[2, 11] == -1.0      # u = 0, 1/2: exact


def _four_dim_condition(process, E, x):
    """Condition numbers of the cancellations inside the Brem / PairProd integrands (all_processes.py:140-206, 532-622)."""
    L = np.longdouble
    E, x1, x2, x3, x4 = [np.asarray(v, dtype=L) for v in (E, x[:, 0], x[:, 1], x[:, 2], x[:, 3])]
    me, pi = L(510.998950e-6), L(np.pi)
    with np.errstate(all="ignore"):
        if process == "PairProd":
            w = E
            e1 = me + x1 * (w - 2 * me); d1 = w / (2 * me) * (x2 + x3); d2 = w / (2 * me) * (x2 - x3); c = np.cos(x4 * 2 * pi)
            e2 = w - e1
            A, B = d1 ** 2 + d2 ** 2, 2 * d1 * d2 * c
            ua, ub = (1 + d1 ** 2) / (2 * e1), (1 + d2 ** 2) / (2 * e2)
            u2 = me ** 2 * (ua + ub) ** 2
            q = A + B + u2
            kq = (A + np.abs(B) + u2) / np.abs(q)
            wu_ku = 0 * kq                                            # u is a sum of positive terms here
            o1, o2 = 1 + d1 ** 2, 1 + d2 ** 2
            T = [-d1 ** 2 / o1 ** 2, -d2 ** 2 / o2 ** 2, w ** 2 / (2 * e1 * e2) * A / (o1 * o2), (e1 / e2 + e2 / e1) * d1 * d2 * c / (o1 * o2)]
        else:
            ml = L(105.6583755e-3) if process == "MuonBrem" else me
            Eg = L(0.001)
            w = Eg + x1 * (E - ml - Eg); d1 = E / (2 * ml) * (x2 + x3); d2 = E / (2 * ml) * (x2 - x3); c = np.cos((x4 - L(0.5)) * 2 * pi)
            e2 = E - w
            A, B = d1 ** 2 + d2 ** 2, 2 * d1 * d2 * c
            ua, ub = (1 + d1 ** 2) / (2 * E), (1 + d2 ** 2) / (2 * e2)
            u2 = ml ** 2 * (ua - ub) ** 2
            q = A - B + u2
            kq = (A + np.abs(B) + u2) / np.abs(q)
            wu_ku = (u2 / np.abs(q)) * (np.abs(ua) + np.abs(ub)) / np.abs(ua - ub)
            o1, o2 = 1 + d1 ** 2, 1 + d2 ** 2
            T = [d1 ** 2 / o1 ** 2, d2 ** 2 / o2 ** 2, w ** 2 / (2 * E * e2) * A / (o1 * o2), -(e2 / E + E / e2) * d1 * d2 * c / (o1 * o2)]
        kT = sum(np.abs(t) for t in T) / np.abs(sum(T))
    f = lambda a: np.nan_to_num(np.asarray(a, dtype=np.float64), nan=1e300, posinf=1e300)
    return f(kq), f(wu_ku), f(kT)


@pytest.mark.parametrize("material", ["graphite", "lead"])
@pytest.mark.parametrize("process", SM)
def test_dsigma_vs_reference_golden(golden, material, process):
    from petite_b200 import _capi as capi
    g = golden("integrands")
    E, x, f = g[f"{material}/{process}/E"], g[f"{material}/{process}/x"], g[f"{material}/{process}/f"]
    sh = shower(material)
    got = probe(sh, capi.PROBE_DSIGMA, CODE[process], np.column_stack([E, x]), 1)[:, 0]
    assert np.array_equal(got == 0, f == 0)
    nz = f != 0
    rel = np.abs(got[nz] - f[nz]) / np.abs(f[nz])
    if process in DIM:
        # the folded form evaluated inside k_sample (probe code + 32) obeys the same bound
        fast = probe(sh, capi.PROBE_DSIGMA, CODE[process] + 32, np.column_stack([E, x]), 1)[:, 0]
        assert np.array_equal(fast == 0, f == 0)
        rel = np.maximum(rel, np.abs(fast[nz] - f[nz]) / np.abs(f[nz]))
        # These integrands subtract nearly equal terms (q^2, the bracket u, T1+T2+T3+T4: all_processes.py:181-201,
        # 572-600), so even the reference's own float64 value carries an amplified rounding error.  Bound: 1e-12 plus
        # machine epsilon times the condition numbers of those three subtractions, evaluated in 80-bit arithmetic.
        kq, wu_ku, kT = _four_dim_condition(process, E, x)
        tol = 1e-12 + 2.3e-16 * (30 * kq + 16 * wu_ku + 8 * kT)[nz]
        assert np.all(rel <= tol), float(np.max(rel / tol))
        assert np.median(rel) < 1e-13
    else:
        assert np.max(rel) < 1e-12


@pytest.mark.parametrize("material", ["graphite", "lead"])
@pytest.mark.parametrize("process,code", [("DarkBrem", 8), ("DarkAnn", 9), ("DarkComp", 10), ("DarkMuonBrem", 11)])
def test_dark_dsigma_vs_reference_golden(golden, material, process, code):
    """Dark integrands (all_processes.py:208-372, 401-466, 625-742 with mV > 0) against values of the unmodified reference
    functions; for the two dark-brem processes also the folded form the sampler evaluates (ds_darkbrem_fast, probe code + 32)."""
    from petite_b200 import _capi as capi
    from tests.test_gpu_dark import dark_shower
    g = golden("integrands")
    E, x, f = g[f"{material}/{process}/E"], g[f"{material}/{process}/x"], g[f"{material}/{process}/f"]
    ds = dark_shower(material, 0.03)
    forms = [probe(ds, capi.PROBE_DSIGMA, code, np.column_stack([E, x]), 1)[:, 0]]
    brem = "Brem" in process
    if brem or process == "DarkAnn":       # DarkAnn: per-sample constants hoisted, powers as exp(b log a) (ds_darkann_c)
        forms.append(probe(ds, capi.PROBE_DSIGMA, code + 32, np.column_stack([E, x]), 1)[:, 0])
    # The dark-brem formula subtracts p^2 + k^2 - 2 p k cos(theta) with 1 - cos(theta) down to 1e-12 (map variable
    # log10(1 - cos)), so even the reference's own float64 value carries a rounding error of about eps / (1 - cos); the
    # GPU forms contract multiply-adds and land within that band (measured: median 5e-12, 7e-6 at 1 - cos = 5e-12, always
    # below 1.1 eps / (1 - cos); bound: 9 eps / (1 - cos)).
    cond = 2e-15 / 10.0 ** x[:, 1] if brem else 0.0
    ann_tol = 0.0
    if process == "DarkAnn":
        # Radiative return: x1 = 1 - (u u_max)^(2 / beta) with 2 / beta ~ 45-90, then the flux factor (1 - x2)^(beta / 2 - 1) with
        # x2 = mV^2 / (x1 s) up to 1 - 1e-12 next to the resonance.  A relative error e_p of the power moves x1 by e_p w (w = 1 - x1)
        # and the result by that times 1 / x1 + 1 / (x1 - y); libm's and libdevice's pow differ by <= 2 ulp, exp(b log a) of the
        # sampler form by <= |b ln a| ulp.  (The reference's own doubles sit up to 2e-5 from an 80-bit evaluation of its formula here;
        # this replaces the flat 1e-9 of round 1.)
        me, al, mV = 510.998950e-6, 1.0 / 137.035999, 0.03
        s_ = 2 * me * (E + me)
        beta = (2 * al / np.pi) * (np.log(s_ / me ** 2) - 1)
        u = x[:, 0] * (1 - mV ** 2 / s_) ** (beta / 2)
        w = u ** (2 / beta)
        x1, y = 1 - w, mV ** 2 / s_
        with np.errstate(all="ignore"):
            e_pow = 2.3e-16 * (4 + (2 / beta) * np.abs(np.log(np.maximum(u, 1e-300))))            # relative error of the power
            dx1 = e_pow * w / np.maximum(x1, 1e-300)                                             # ... of x1
            near = np.maximum(x1, 1e-300) / np.maximum(x1 - y, 1e-300)                           # 1 / (1 - x2)
            ann_tol = 1e-12 + 4 * (dx1 * (2 + near) + 2 * 2.3e-16 * near)
            ann_tol = np.where(np.isfinite(ann_tol), ann_tol, 1.0)
    for got in forms:
        assert np.array_equal(got == 0, f == 0)
        for Einc in np.unique(E):
            sel = E == Einc
            scale = np.max(np.abs(f[sel])) if np.any(f[sel] != 0) else 1.0
            rtol = ann_tol[sel] if process == "DarkAnn" else 1e-12
            tol = (rtol + (1e-11 + cond[sel] if brem else 0.0)) * np.abs(f[sel]) + (1e-9 * scale if brem else 0.0)
            worst = float(np.max(np.abs(got[sel] - f[sel]) / (tol + 1e-300)))
            assert worst <= 1.0, (process, Einc, worst)
    if brem:       # the folded form follows the plain one far more closely than either follows the reference
        a, b = forms
        nz = a != 0
        assert np.all(np.abs(a[nz] - b[nz]) <= (1e-10 + cond[nz]) * np.abs(a[nz]))
        assert np.median(np.abs(a[nz] - b[nz]) / np.abs(a[nz])) < 1e-12


@pytest.mark.parametrize("material", ["graphite", "lead"])
def test_nsigma_vs_reference_golden(golden, material):
    from petite_b200 import _capi as capi
    g = golden("nsigma")
    sh = shower(material)
    E = g[f"{material}/E"]
    for P in SM:
        got = probe(sh, capi.PROBE_NSIGMA, CODE[P], E[:, None], 1)[:, 0]
        want = g[f"{material}/{P}"]
        assert np.all(np.abs(got - want) <= 1e-12 * np.abs(want)), P
        # host twin used for the reference-compatible _NSigma* attributes
        assert np.all(np.abs(sh._nsigma_tables[P](E) - want) <= 1e-12 * np.abs(want)), P
    assert np.allclose([sh._minimum_calculable_energy[k] for k in (11, -11, 22, 13, -13)], g[f"{material}/min_calc"], rtol=0, atol=0)
    for pid in (22, 11, -11, 13):
        got = np.array([sh.get_mfp([pid, e]) for e in E[::7]])
        assert np.all(np.abs(got - g[f"{material}/mfp/{pid}"][::7]) <= 1e-12 * got)


def test_map_transform_vs_oracle():
    from petite_b200 import _capi as capi
    from oracle.vegasmap import map_points
    from oracle.findmax import split_grid
    sh = shower()
    rng = np.random.default_rng(3)
    for P in SM:
        ms = sh._maps[P]
        for ie in (0, 37, 99):
            y = rng.random((256, ms.dim))
            y[0] = 0.0
            y[1] = 1.0 - 2.0 ** -53
            got = probe(sh, capi.PROBE_MAP, CODE[P], np.column_stack([np.full(len(y), ie), y]), ms.dim + 1)
            x, jac = map_points(split_grid(ms.grid[ie], ms.ninc), y)
            assert np.array_equal(got[:, :ms.dim], x)                      # same roundings -> bit-exact
            assert np.max(np.abs(got[:, ms.dim] - jac) / jac) < 1e-14


def test_kinematics_vs_reference_golden(golden):
    from petite_b200 import _capi as capi
    g = golden("kinematics")
    sh = shower()
    for P in SM + ["SMDecay"]:
        code = CODE.get(P, 12)
        got = probe(sh, capi.PROBE_KIN, code, g[f"{P}/in"], 8)
        want = g[f"{P}/out"]
        scale = np.max(np.abs(want), axis=1, keepdims=True)
        assert np.max(np.abs(got - want) / scale) < 1e-12, P


@pytest.mark.parametrize("material", ["graphite", "lead"])
def test_multiple_scattering_vs_reference_golden(golden, material):
    """theta0 contains p = m beta / sqrt(1 - beta^2): a last-bit difference in |p| is amplified by gamma^2, so the
    bound is 1e-12 |p| plus that conditioning term (SURVEY.md hard part 7: mixed abs/rel at cancellation points)."""
    from petite_b200 import _capi as capi
    g = golden("mcs")
    sh = shower(material)
    Z = 6 if material == "graphite" else 82
    sel = g["inp"][:, 10] == Z
    a, want = g["inp"][sel][:, :10], g["out"][sel]
    got = probe(sh, capi.PROBE_MCS, 0, a, 4)
    pn = np.linalg.norm(a[:, 1:4], axis=1)
    gamma2 = (a[:, 0] / a[:, 5]) ** 2
    dtheta = np.linalg.norm(np.cross(want[:, 1:], a[:, 1:4]), axis=1) / pn ** 2          # scattering angle
    tol = 1e-12 * pn + 4 * 2.3e-16 * gamma2 * dtheta * pn
    assert np.all(np.max(np.abs(got - want), axis=1) <= tol)
    assert np.array_equal(got[:, 0], want[:, 0])                                           # energy untouched


@pytest.mark.parametrize("material", ["graphite", "lead"])
def test_mcs_fast_vs_reference_golden(golden, material):
    """The folded multiple-scattering form the sub-step loop runs (mcs_fast: 2 divisions, constant-memory log / sincos kernels,
    branch-free reciprocals) against moliere.get_scattered_momentum_fast of the unmodified reference - same golden vectors and
    the same condition-number bound as the plain form above (the reference's own value carries the gamma^2-amplified rounding
    of m beta / sqrt(1 - beta^2); mcs_fast itself never forms that difference)."""
    from petite_b200 import _capi as capi
    g = golden("mcs")
    sh = shower(material)
    Z = 6 if material == "graphite" else 82
    sel = g["inp"][:, 10] == Z
    a, want = g["inp"][sel][:, :10], g["out"][sel]
    got = probe(sh, capi.PROBE_MCS_FAST, 0, a, 4)
    plain = probe(sh, capi.PROBE_MCS, 0, a, 4)
    pn = np.linalg.norm(a[:, 1:4], axis=1)
    gamma2 = (a[:, 0] / a[:, 5]) ** 2
    dtheta = np.linalg.norm(np.cross(want[:, 1:], a[:, 1:4]), axis=1) / pn ** 2
    tol = 1e-12 * pn + 4 * 2.3e-16 * gamma2 * dtheta * pn
    assert np.all(np.max(np.abs(got - want), axis=1) <= tol)
    assert np.array_equal(got[:, 0], want[:, 0])
    # against the plain GPU form (no gamma^2 term on either side): re-association only
    worst = float(np.max(np.max(np.abs(got - plain), axis=1) / pn))
    print("mcs_fast vs mcs_apply, max |dp| / |p| =", worst)
    assert worst < 2e-12          # measured 3e-14 (graphite) / 1.6e-13 (lead: (1 + v) log(1 + v) / v - 1 cancels for thin steps)


@pytest.mark.parametrize("mV", [0.003, 0.03])
def test_dark_kinematics_vs_reference_golden(golden, mV):
    """kin_darkbrem_V / kin_darkann_V / kin_compton_bound_V against l_to_lV_fourvecs, radiative_return_fourvecs and
    compton_fourvecs_boundelectron of the unmodified reference (kinematics.py:43-68, 267-299, 134-183).  Bound: 1e-12 of the
    vector's largest component; the radiative-return x1 = 1 - (u umax)^(2 / beta) (2 / beta ~ 40) turns a 2-ulp difference
    between libdevice's and libm's pow into 4e-16 / x1, the bound-electron boost 1 / sqrt(1 - b0^2) amplifies by gamma0^2."""
    from petite_b200 import _capi as capi
    from petite_b200.tables import mv_tag
    from tests.test_gpu_dark import dark_shower
    g = golden("dark_kinematics")
    ds = dark_shower("graphite", mV)
    me, alpha = 510.998950e-6, 1.0 / 137.035999
    for P, code in (("DarkBrem", 8), ("DarkMuonBrem", 11), ("DarkAnn", 9), ("DarkComp", 10)):
        a, want = g[f"{mv_tag(mV)}/{P}/in"], g[f"{mv_tag(mV)}/{P}/out"]
        got = probe(ds, capi.PROBE_DARKKIN, code, a, 4)
        scale = np.max(np.abs(want), axis=1)
        err = np.max(np.abs(got - want), axis=1) / scale
        tol = np.full(len(a), 1e-12)
        if P == "DarkAnn":
            s = 2 * me * (me + a[:, 0])
            beta = (2 * alpha / np.pi) * (np.log(s / me ** 2) - 1)
            x1 = 1 - (a[:, 2] * (1 - mV ** 2 / s) ** (beta / 2)) ** (2 / beta)
            # ... and the boost back to the lab forms sqrt(p0^2 - p3^2) = m_e from two numbers of size s / 4 (radiative_return.py:18-24):
            # conditioning Ee / (2 m_e); the reference's own doubles sit up to 3 of these units from an 80-bit evaluation
            tol = tol + 2e-15 / x1 + 8 * 2.3e-16 * a[:, 0] / (2 * me)
        if P == "DarkComp":
            Eg, Pe, cte = a[:, 0], a[:, 8], a[:, 9]
            b0 = np.sqrt(Eg ** 2 + 2 * cte * Eg * Pe + Pe ** 2) / (Eg + np.sqrt(me ** 2 + Pe ** 2))
            # ... and the photon-electron system is rotated back with sin = sqrt(1 - ctz^2), ctz = (Eg + cte Pe) / |p_tot| within 1e-14
            # of 1 for a GeV photon on a keV electron: a half-ulp of ctz moves the sine by 1.1e-16 / sin (the reference's own
            # doubles sit up to 1e-9 from an 80-bit evaluation of its formula there)
            ctz = (Eg + cte * Pe) / np.sqrt(Eg ** 2 + 2 * cte * Eg * Pe + Pe ** 2)
            tol = tol + 8 * 2.3e-16 / (1 - b0 ** 2) + 8 * 2.3e-16 / np.sqrt(np.maximum(1 - ctz ** 2, 1e-30))
        print(P, mV, "max err / tol", float(np.max(err / tol)), "max err", float(np.max(err)))
        assert np.all(err <= tol), (P, float(np.max(err / tol)))


@pytest.mark.parametrize("material", ["graphite", "lead"])
def test_substep_vs_oracle(material):
    """ONE iteration of the dE/dx + multiple-scattering loop exactly as k_loop runs it (substep(): per-species summed n*sigma
    table with a hint, fast_rcp, hot_exp_neg_step, carried |p|, mcs_fast) against the oracle's statement of
    shower.py:559-581 (libm, term-by-term n*sigma sums, lose_energy + get_scattered_momentum_fast) on 24 000 random tracks:
    the loop decision (hard scatter / below threshold / step) must be IDENTICAL, the state within the multiple-scattering
    condition bound.  Reports the margin of the closest hard-scatter decision."""
    import math
    from petite_b200 import _capi as capi
    from oracle.shower import OracleShower
    from oracle.draws import CounterDraws
    from oracle import physics as phy, consts as OC
    sh = shower(material)
    o = OracleShower(None, material, 0.010, rng="counter")
    rng = np.random.default_rng(77)
    n = 8000
    rows, exp = [], []
    for pid in (11, -11, 13):
        m = OC.MASS[pid]
        pmin = max(o.min_calc[pid], o.min_energy, m)
        E = np.where(rng.random(n) < 0.05, pmin * rng.uniform(0.5, 1.0, n), pmin * 10 ** rng.uniform(0, np.log10(100.0 / pmin), n))
        if pid == 13:
            E = np.maximum(E, m * 1.0001)
        d = rng.normal(size=(n, 3))
        small = rng.random(n) < 0.5                      # shower particles are collimated along z
        d[small, :2] *= 1e-3
        d /= np.linalg.norm(d, axis=1, keepdims=True)
        r = rng.normal(size=(n, 3))
        keys = rng.integers(0, 2 ** 32, size=(n, 2))
        it = rng.integers(0, 60, n)
        for k in range(n):
            pn = math.sqrt(max(E[k] ** 2 - m ** 2, 0.0))
            p4 = [float(E[k]), pn * d[k, 0], pn * d[k, 1], pn * d[k, 2]]
            rows.append([pid] + p4 + list(r[k]) + [float(keys[k, 0]), float(keys[k, 1]), float(it[k]), 1.0])
            dr = CounterDraws((int(keys[k, 0]), int(keys[k, 1])))
            if not p4[0] >= pmin:
                exp.append([1.0] + p4 + list(r[k]) + [0.0, 0.0, 1.0])
                continue
            mfp = o.get_mfp(pid, p4[0])
            u_hard, u_dz = dr.substep(int(it[k]))
            dz = mfp / (6 + 14 * u_dz)
            margin = abs(u_hard - math.exp(-dz / mfp))
            if u_hard > math.exp(-dz / mfp):
                exp.append([1.0] + p4 + list(r[k]) + [dz, margin, 1.0])
                continue
            pf = phy.lose_energy(p4, m, o.dEdx * 0.1 * dz)
            pnf = phy.norm3(pf[1:])
            rf = [r[k, j] + pf[1 + j] / pnf * dz for j in range(3)] if pnf > 0 else list(r[k])
            pf = o._mcs(pf, dz, m, dr, int(it[k]))
            exp.append([0.0] + list(pf) + rf + [dz, margin, (p4[0] / m) ** 2])
    a, exp = np.array(rows), np.array(exp)
    got = probe(sh, capi.PROBE_SUBSTEP, 0, a, 10)
    flips = int(np.sum(got[:, 0] != exp[:, 0]))
    live = exp[:, 9] > 0
    print(material, "sub-steps", len(a), "decision flips", flips, "closest hard-scatter margin", float(np.min(exp[live, 9])))
    assert flips == 0
    st = exp[:, 0] == 0
    pn = np.linalg.norm(exp[st, 2:5], axis=1)
    assert np.all(np.abs(got[st, 1] - exp[st, 1]) <= 1e-13 * exp[st, 1])                       # energy
    assert np.all(np.abs(got[st, 8] - exp[st, 8]) <= 1e-13 * exp[st, 8])                       # delta_z
    assert np.all(np.max(np.abs(got[st, 5:8] - exp[st, 5:8]), axis=1) <= 1e-13 * np.maximum(1.0, np.max(np.abs(exp[st, 5:8]), axis=1)))
    dth = np.linalg.norm(np.cross(got[st, 2:5], a[st, 2:5]), axis=1) / np.maximum(pn ** 2, 1e-300)     # stopped tracks: |p| = 0
    tol = 1e-12 * pn + 4 * 2.3e-16 * exp[st, 10] * dth * pn
    assert np.all(np.max(np.abs(got[st, 2:5] - exp[st, 2:5]), axis=1) <= tol + 1e-300)
    assert np.array_equal(got[st, 9], a[st, 10] + 1)
