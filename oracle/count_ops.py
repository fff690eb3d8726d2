"""Counted fp64 operations of the reference formulas (oracle restatement), per unit of work.  TEST INFRASTRUCTURE.

    python -m oracle.count_ops          # prints the constants that petite_b200/roofline.py carries

SURVEY.md 8(d): "fix the F constants once by counting operations in the oracle (div, sqrt = 1 flop; cos/sin/exp/log/pow/atan2/
acos = 1 'special' counted separately)".  Two instruments, no change to the oracle's code:
  * the vectorised integrands (oracle/integrands.py) run on a NumPy ndarray subclass whose __array_ufunc__ tallies every
    elementwise ufunc call by name;
  * the scalar pieces (sub-step body of shower.py:559-581, kinematics + rotation, map transform) run on a float subclass that
    tallies its arithmetic, with the ``math`` module seen by oracle.physics / oracle.shower replaced by a tallying proxy.
Comparisons, selections and integer index arithmetic are not counted.
"""
import math
from collections import Counter

import numpy as np

ARITH = {"add", "subtract", "multiply", "divide", "true_divide", "sqrt", "square", "negative", "absolute", "reciprocal"}
SPECIAL = {"cos", "sin", "tan", "exp", "exp2", "log", "log10", "power", "float_power", "arctan", "arctan2", "arccos", "arcsin", "sinh", "cosh"}
FREE = {"negative", "absolute"}       # sign manipulations: no pipe instruction of their own


class CountArr(np.ndarray):
    tally = Counter()

    def __array_ufunc__(self, ufunc, method, *inputs, **kw):
        args = [np.asarray(a) if isinstance(a, CountArr) else a for a in inputs]
        if "out" in kw:
            kw["out"] = tuple(np.asarray(o) if isinstance(o, CountArr) else o for o in kw["out"])
        res = getattr(ufunc, method)(*args, **kw)
        n = int(np.size(res)) if not isinstance(res, tuple) else int(np.size(res[0]))
        name = ufunc.__name__
        if name in ("power", "float_power") and np.isscalar(args[1]) and float(args[1]).is_integer():
            k = int(args[1])                       # x ** k with a small integer k: |k| - 1 multiplies (+ one division if k < 0)
            CountArr.tally["multiply"] += n * max(abs(k) - 1, 0)
            if k < 0:
                CountArr.tally["divide"] += n
        else:
            CountArr.tally[name] += n
        if isinstance(res, np.ndarray) and res.dtype.kind == "f":
            return res.view(CountArr)
        return res


class CF(float):
    tally = Counter()

    def _b(name):
        def f(self, o):
            CF.tally[name] += 1
            return CF(getattr(float, "__%s__" % name)(float(self), float(o)))
        return f
    __add__ = _b("add"); __radd__ = _b("radd"); __sub__ = _b("sub"); __rsub__ = _b("rsub")
    __mul__ = _b("mul"); __rmul__ = _b("rmul"); __truediv__ = _b("truediv"); __rtruediv__ = _b("rtruediv")

    def __pow__(self, o):
        if float(o) == 2.0:
            CF.tally["mul"] += 1
        else:
            CF.tally["special:pow"] += 1
        return CF(float(self) ** float(o))

    def __rpow__(self, o):
        CF.tally["special:pow"] += 1
        return CF(float(o) ** float(self))

    def __neg__(self):
        return CF(-float(self))

    def __abs__(self):
        return CF(abs(float(self)))


class MathProxy:
    pi, e, inf = math.pi, math.e, math.inf

    def __getattr__(self, name):
        fn = getattr(math, name)

        def f(*a):
            if name == "sqrt":
                CF.tally["sqrt"] += 1
            elif name in ("floor", "ceil", "isnan", "isinf", "copysign", "fabs"):
                pass
            else:
                CF.tally["special:" + name] += 1
            r = fn(*[float(x) for x in a])
            return CF(r) if isinstance(r, float) else r
        return f


def split(tally):
    flops = sum(v for k, v in tally.items() if (k in ARITH and k not in FREE) or k in ("add", "radd", "sub", "rsub", "mul", "rmul", "truediv", "rtruediv", "sqrt"))
    special = sum(v for k, v in tally.items() if k in SPECIAL or k.startswith("special:"))
    return flops, special


def count_integrands(n_points=64):
    """Per-trial flops / specials of every integrand at points drawn through the shipped maps (row 60), including the map
    transform (oracle/vegasmap.py) and the accept test (one multiply on each side)."""
    from . import integrands as I
    from .consts import SM_PROCESSES, DARK_PROCESSES, TARGETS, m_electron, m_muon
    from .findmax import split_grid
    from .vegasmap import map_points
    from .shower import DATA_DIR
    out = {}
    rng = np.random.default_rng(1)
    t = TARGETS["lead"]
    sm = np.load(DATA_DIR + "sm_maps.npz"); dk = np.load(DATA_DIR + "dark_maps_mV0.03.npz")
    for P in SM_PROCESSES + DARK_PROCESSES:
        z = sm if P in SM_PROCESSES else dk
        grid = split_grid(z[f"{P}/grid"][60], z[f"{P}/ninc"])
        E = float(z[f"{P}/E"][60])
        y = rng.random((n_points, len(grid)))
        CountArr.tally = Counter()
        x, jac = map_points([g.view(CountArr) for g in grid], y.view(CountArr))
        ev = dict(E_inc=E, Z_T=t["Z_T"], A_T=t["A_T"], mT=t["A_T"], mV=0.0 if P in SM_PROCESSES else 0.03, Eg_min=0.001, Ee_min=0.005,
                  m_lepton=m_muon if "Muon" in P else m_electron)
        cols, I._cols = I._cols, (lambda a, n: [a[..., i] for i in range(n)])        # np.asarray would drop the counting subclass
        try:
            I.DSIGMA[P](np.asarray(x).view(CountArr), ev)
        finally:
            I._cols = cols
        fl, sp = split(CountArr.tally)
        out[P] = (fl / n_points + 3, sp / n_points, dict(CountArr.tally))     # + accept test: jac/B, * f, max_F * u
    return out


def count_scalar():
    """One dE/dx + multiple-scattering sub-step (shower.py:559-581 as oracle.shower.OracleShower.propagate states it) and one
    hard-scatter kinematics + rotation to the lab frame."""
    from . import physics as phy, shower as shw
    from .shower import OracleShower
    from .draws import CounterDraws
    o = OracleShower(None, "lead", 0.010, rng="counter")
    proxy = MathProxy()
    old = phy.math, shw.math
    phy.math = shw.math = proxy
    res = {}
    try:
        # sub-step: get_mfp (3 table interpolations for e+: slope, product, sum each) + body
        CF.tally = Counter()
        E = CF(1.234)
        m = 510.998950e-6
        p4 = [E, CF(0.01), CF(-0.02), CF(math.sqrt(1.234 ** 2 - m * m - 0.0005))]
        ns_terms = 3                                                   # Brem + Bhabha + Ann
        CF.tally["table"] += 0
        mfp = CF(o.get_mfp(-11, float(E)))
        for _ in range(ns_terms):                                      # scipy interp1d linear: (y1-y0)/(x1-x0)*(x-x0)+y0
            CF.tally["sub"] += 3; CF.tally["truediv"] += 1; CF.tally["mul"] += 1; CF.tally["add"] += 1
        CF.tally["add"] += ns_terms - 1; CF.tally["truediv"] += 1       # sum, cmtom / ns
        u_hard, u_dz = CF(0.3), CF(0.6)
        dz = mfp / (6 + (20 - 6) * u_dz)
        proxy.exp(-dz / mfp)
        pf = phy.lose_energy(p4, CF(m), CF(o.dEdx * 0.1) * dz)
        pn = phy.norm3(pf[1:])
        rf = [CF(0.0) + pf[1 + k] / pn * dz for k in range(3)]
        CF.tally["special:log"] += 1; CF.tally["sqrt"] += 1; CF.tally["special:cos"] += 1; CF.tally["special:sin"] += 1   # two normals (Box-Muller)
        pn2 = CF(0.0) + pf[1] * pf[1] + pf[2] * pf[2] + pf[3] * pf[3]; proxy.sqrt(pn2)    # np.linalg.norm inside mcs_scatter
        phy.mcs_scatter([CF(v) for v in pf], CF(o.rho) * (dz / CF(0.01)), CF(o.A), CF(o.Z), CF(1.0), CF(m), 1, CF(0.3), CF(-1.2), CF(0.77))
        res["substep"] = split(CF.tally) + (dict(CF.tally),)
        # kinematics (brem) + rotation matrix + two rotations
        CF.tally = Counter()
        v1, v2 = phy.kin_brem(CF(1.234), CF(m), [CF(0.3), CF(1e-3), CF(4e-4), CF(0.6)], CF(0.41))
        R = phy.rotation_matrix([CF(v) for v in p4])
        phy.rotate(R, v1[1:]); phy.rotate(R, v2[1:])
        res["kinematics"] = split(CF.tally) + (dict(CF.tally),)
    finally:
        phy.math, shw.math = old
    return res


if __name__ == "__main__":
    integ = count_integrands()
    print("FLOPS_TRIAL = {" + ", ".join(f'"{p}": {round(v[0])}' for p, v in integ.items()) + "}")
    print("SPECIAL_TRIAL = {" + ", ".join(f'"{p}": {round(v[1])}' for p, v in integ.items()) + "}")
    for k, v in count_scalar().items():
        print(k, "flops", v[0], "specials", v[1], v[2])
