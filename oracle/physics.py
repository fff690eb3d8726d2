"""Per-particle physics pieces (oracle).  TEST INFRASTRUCTURE - see oracle/__init__.py.

Scalar restatements of reference ``particle.py`` (lose_energy :143, rotation_matrix :176, boost_matrix :187,
two_body_decay :209), ``moliere.py`` (:196-219, :265-400) and ``kinematics.py``.  Random inputs are passed
in explicitly (uniforms / standard normals) so the same functions serve stream mode and counter mode.
"""
import math

import numpy as np

from .consts import alpha_em, m_electron, m_muon, MeV

TWO_PI = 2.0 * math.pi
EGAMMA_MIN_KIN = 0.001  # kinematics.py:9 module constant (Q-6)


def norm3(v):
    return math.sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2])


def lose_energy(pf, mass, value):
    """particle.py:143-153.  pf: [E,px,py,pz] -> new list."""
    E0, px, py, pz = pf
    p30 = norm3((px, py, pz))
    Eu = E0 - value
    if Eu <= mass:
        Eu = mass
    p3f = math.sqrt(Eu * Eu - mass * mass)
    if p3f > 0.0:
        return [Eu, px / p30 * p3f, py / p30 * p3f, pz / p30 * p3f]
    return [mass, 0.0, 0.0, 0.0]


def rotation_matrix(pf):
    """particle.py:176-185: Rz(phi) Ry(theta) taking z-hat to the direction of pf."""
    _, px, py, pz = pf
    th = math.acos(pz / math.sqrt(px * px + py * py + pz * pz))
    ph = math.atan2(py, px)
    ct, st, cp, sp = math.cos(th), math.sin(th), math.cos(ph), math.sin(ph)
    return [[ct * cp, -sp, st * cp], [ct * sp, cp, st * sp], [-st, 0.0, ct]]


def rotate(R, v):
    return [R[0][0] * v[0] + R[0][1] * v[1] + R[0][2] * v[2],
            R[1][0] * v[0] + R[1][1] * v[1] + R[1][2] * v[2],
            R[2][0] * v[0] + R[2][1] * v[1] + R[2][2] * v[2]]


# ---------------- multiple Coulomb scattering (fast / Lynch-Dahl) ----------------

def mcs_theta0(t, beta, A, Z, z=1.0, m_lepton=m_electron):
    """moliere.py:196-219, 265-281: Lynch & Dahl width (F = 0.98)."""
    F = 0.98
    p = (m_lepton / MeV) * beta / math.sqrt(1.0 - beta * beta)
    chic2 = 0.157 * Z * (Z + 1) * (t / A) * (z / (p * beta)) ** 2
    chia2 = 2.007e-5 * Z ** (2.0 / 3.0) * (1.0 + 3.34 * (Z * z * alpha_em / beta) ** 2) / (p * p)
    omega = chic2 / chia2
    v = 0.5 * omega / (1.0 - F)
    return math.sqrt(chic2 * ((1.0 + v) * math.log(1.0 + v) / v - 1) / (1.0 + F * F))


def align_to_z_matrix(v):
    """moliere.py:287-348 (get_rotation_matrix): R with R v = |v| z-hat, incl. the duplicated branch (Q-13)."""
    vx, vy, vz = v
    if abs(vx) > 0.0 and abs(vy) > 0.0:
        a = math.atan(abs(vy / vx))
        if vx > 0.0 and vy > 0.0:
            a = -a
        if vx < 0.0 and vy > 0.0:
            a = -(math.pi - a)
        if vx < 0.0 and vy < 0.0:
            a = -(math.pi + a)
        if vx > 0.0 and vy < 0.0:
            a = -(2.0 * math.pi - a)
        ca, sa = math.cos(a), math.sin(a)
    elif abs(vy) > 0.0:
        ca, sa = 0.0, 1.0
    else:
        ca, sa = 1.0, 0.0
    vxp = vx * ca - vy * sa
    if abs(vz) > 0.0 and abs(vxp) > 0.0:
        b = math.atan(abs(vxp / vz))
        if vz > 0.0 and vxp > 0.0:
            b = -b
        if vz < 0.0 and vxp > 0.0:
            b = -(math.pi - b)
        if vz < 0.0 and vxp < 0.0:
            b = -(math.pi + b)
        if vz > 0.0 and vxp > 0.0:       # duplicated condition in the reference; acts on the already negated b
            b = -(2.0 * math.pi - b)
        cb, sb = math.cos(b), math.sin(b)
    elif vxp > 0.0:
        cb, sb = 0.0, -1.0
    elif vxp < 0.0:
        cb, sb = 0.0, 1.0
    else:
        cb, sb = 1.0, 0.0
    # Rb @ Ra
    return [[cb * ca, -cb * sa, sb], [sa, ca, 0.0], [-sb * ca, sb * sa, cb]]


def mcs_scatter(p4, t, A, Z, rescale, m_lepton, sign, z1, z2, u_phi):
    """moliere.py:350-400 (get_scattered_momentum_fast).

    ``sign`` in {-1,+1}, ``z1, z2`` standard normals, ``u_phi`` uniform in [0,1).
    """
    p3 = p4[1:]
    # np.linalg.norm as in the reference: beta feeds sqrt(1 - beta^2), which amplifies a last-bit difference of the
    # norm by gamma^2 (up to ~1e8 for 10 GeV electrons), so the summation order matters here
    pn = float(np.linalg.norm(np.asarray(p3, dtype=np.float64)))
    if not pn > 0:
        return list(p4)
    beta = pn / p4[0]
    Rinv = align_to_z_matrix(p3)
    th0 = mcs_theta0(t, beta, A, Z, 1.0, m_lepton)
    g1 = 0.0 + z1 * th0
    g2 = 0.0 + z2 * th0
    theta = sign * math.sqrt(g1 * g1 + g2 * g2) * rescale
    phi = 0.0 + (TWO_PI - 0.0) * u_phi
    cth, sth, cph, sph = math.cos(theta), math.sin(theta), math.cos(phi), math.sin(phi)
    # Rphi @ Rtheta @ z-hat = (sph*sth, -cph*sth, cth)
    q = [pn * (sph * sth), pn * (-cph * sth), pn * cth]
    # back to the lab with the transpose
    out = [Rinv[0][0] * q[0] + Rinv[1][0] * q[1] + Rinv[2][0] * q[2],
           Rinv[0][1] * q[0] + Rinv[1][1] * q[1] + Rinv[2][1] * q[2],
           Rinv[0][2] * q[0] + Rinv[1][2] * q[1] + Rinv[2][2] * q[2]]
    return [p4[0]] + out


def normals_from_uniforms(u_angle, u_radius):
    """CPython random.gauss: z1 = cos(2 pi u1) * sqrt(-2 ln(1-u2)), cached z2 = sin(...) * same."""
    x2pi = u_angle * TWO_PI
    g2rad = math.sqrt(-2.0 * math.log(1.0 - u_radius))
    return math.cos(x2pi) * g2rad, math.sin(x2pi) * g2rad


# ---------------- kinematics: sampled variables -> two four-vectors (parent along z) ----------------

def kin_brem(E, m_lepton, x, u_az):
    """kinematics.py:10-41 (e_to_egamma_fourvecs); returns [lepton, photon]."""
    x1, x2, x3, x4 = x[:4]
    ep = E
    w = EGAMMA_MIN_KIN + x1 * (ep - m_lepton - EGAMMA_MIN_KIN)
    ct = math.cos((x2 + x3) / 2)
    ctp = math.cos((x2 - x3) * ep / (2 * (ep - w)))
    ph = (x4 - 1 / 2) * 2.0 * math.pi
    epp = ep - w
    pp = _sqrt(epp ** 2 - m_lepton ** 2)
    al = u_az * TWO_PI
    cal, sal = math.cos(al), math.sin(al)
    st, stp = _sqrt(1.0 - ct ** 2), _sqrt(1.0 - ctp ** 2)
    sp, cp = math.sin(ph), math.cos(ph)
    g = [w, w * cal * st, w * sal * st, w * ct]
    l = [epp, pp * (sal * sp * stp + cal * (ctp * st - cp * ct * stp)),
         pp * (ctp * sal * st - (cp * ct * sal + cal * sp) * stp), pp * (ct * ctp + cp * st * stp)]
    return [l, g]


def _sqrt(v):
    return math.sqrt(v) if v >= 0 else float("nan")


def kin_pairprod(E, x, u_az):
    """kinematics.py:70-102 (gamma_to_epem_fourvecs); returns [positron, electron]."""
    w = E
    x1, x2, x3, x4 = x[:4]
    me = m_electron
    epp = me + x1 * (w - 2 * me)
    ctp = math.cos(w * (x2 + x3) / (2 * epp))
    ctm = math.cos(w * (x2 - x3) / (2 * (w - epp)))
    ph = x4 * 2 * math.pi
    epm = w - epp
    pm, pp = _sqrt(epm ** 2 - me ** 2), _sqrt(epp ** 2 - me ** 2)
    al = u_az * TWO_PI
    cal, sal = math.cos(al), math.sin(al)
    stp, stm = _sqrt(1.0 - ctp ** 2), _sqrt(1.0 - ctm ** 2)
    spal, cpal = math.sin(ph + al), math.cos(ph + al)
    return [[epp, pp * stp * cal, pp * stp * sal, pp * ctp], [epm, pm * stm * cpal, pm * stm * spal, pm * ctm]]


def kin_compton(E, x, u_az, mV=0.0):
    """kinematics.py:104-132 (compton_fourvecs); returns [electron, photon/V]."""
    me = m_electron
    Eg, ct = E, x[0]
    s = me ** 2 + 2 * Eg * me
    rs = math.sqrt(s)
    Ee0 = (s + me ** 2) / (2.0 * rs)
    Ee = (s - mV ** 2 + me ** 2) / (2 * rs)
    EV = (s + mV ** 2 - me ** 2) / (2 * rs)
    pF = _sqrt(Ee ** 2 - me ** 2)
    g0 = Ee0 / me
    b0 = 1.0 / g0 * math.sqrt(g0 ** 2 - 1.0)
    ph = u_az * TWO_PI
    st = _sqrt(1 - ct ** 2)
    sp, cp = math.sin(ph), math.cos(ph)
    pe = [g0 * Ee + b0 * g0 * pF * ct, -pF * st * sp, -pF * st * cp, b0 * g0 * Ee + g0 * pF * ct]
    pV = [g0 * EV - b0 * g0 * pF * ct, pF * st * sp, pF * st * cp, b0 * g0 * EV - g0 * pF * ct]
    return [pe, pV]


def kin_annihilation(E, x, u_az, mV=0.0):
    """kinematics.py:301-334 (annihilation_fourvecs); returns [photon, photon/V]."""
    me = m_electron
    Ee, ct = E, x[0]
    s = 2 * me * (Ee + me)
    rs = math.sqrt(s)
    EeCM = rs / 2.0
    Eg = (s - mV ** 2) / (2 * rs)
    EV = (s + mV ** 2) / (2 * rs)
    pF = Eg
    g0 = EeCM / me
    b0 = 1.0 / g0 * math.sqrt(g0 ** 2 - 1.0)
    ph = u_az * TWO_PI
    st = _sqrt(1 - ct ** 2)
    sp, cp = math.sin(ph), math.cos(ph)
    pg = [g0 * Eg - b0 * g0 * pF * ct, -pF * st * sp, -pF * st * cp, b0 * g0 * Eg - g0 * pF * ct]
    pV = [g0 * EV + b0 * g0 * pF * ct, pF * st * sp, pF * st * cp, b0 * g0 * EV + g0 * pF * ct]
    return [pg, pV]


def kin_ee(E, x, u_az):
    """kinematics.py:213-237 (ee_to_ee_fourvecs): Moller / Bhabha; returns [scattered, struck electron]."""
    me = m_electron
    ct = x[0]
    s = 2 * me ** 2 + 2 * E * me
    Ee0 = math.sqrt(s) / 2.0
    pF = math.sqrt(Ee0 ** 2 - me ** 2)
    g0 = Ee0 / me
    b0 = 1.0 / g0 * math.sqrt(g0 ** 2 - 1.0)
    ph = u_az * TWO_PI
    st = _sqrt(1 - ct ** 2)
    sp, cp = math.sin(ph), math.cos(ph)
    a = [g0 * Ee0 + b0 * g0 * pF * ct, -pF * st * sp, -pF * st * cp, b0 * g0 * Ee0 + g0 * pF * ct]
    b = [g0 * Ee0 - b0 * g0 * pF * ct, pF * st * sp, pF * st * cp, b0 * g0 * Ee0 - g0 * pF * ct]
    return [a, b]


def kin_mue(E, x, u_az):
    """kinematics.py:239-265 (mue_to_mue_fourvecs); returns [muon, electron]."""
    me, mm = m_electron, m_muon
    ct = x[0]
    s = me ** 2 + mm ** 2 + 2 * E * me
    rs = math.sqrt(s)
    Ee0 = (s + me ** 2 - mm ** 2) / (2.0 * rs)
    Em0 = (s + mm ** 2 - me ** 2) / (2.0 * rs)
    pe = math.sqrt(Ee0 ** 2 - me ** 2)
    pm = math.sqrt(Em0 ** 2 - mm ** 2)
    g0 = Ee0 / me
    b0 = 1.0 / g0 * math.sqrt(g0 ** 2 - 1.0)
    ph = u_az * TWO_PI
    st = _sqrt(1 - ct ** 2)
    sp, cp = math.sin(ph), math.cos(ph)
    mu = [g0 * Em0 + b0 * g0 * pm * ct, pm * st * sp, pm * st * cp, b0 * g0 * Em0 + g0 * pm * ct]
    el = [g0 * Ee0 - b0 * g0 * pe * ct, -pe * st * sp, -pe * st * cp, b0 * g0 * Ee0 - g0 * pe * ct]
    return [mu, el]


def kin_darkbrem(E, m_lepton, x, u_az, mV):
    """kinematics.py:43-68 (l_to_lV_fourvecs); returns [lepton (unchanged beam), V]."""
    ep = E
    w = x[0] * ep
    ct = 1 - 10 ** x[1]
    p, k = _sqrt(ep ** 2 - m_lepton ** 2), _sqrt(w ** 2 - mV ** 2)
    al = u_az * TWO_PI
    cal, sal = math.cos(al), math.sin(al)
    st = _sqrt(1.0 - ct ** 2)
    return [[ep, 0.0, 0.0, p], [w, k * cal * st, k * sal * st, k * ct]]


def kin_darkann(E, x, mV):
    """kinematics.py:267-299 (radiative_return_fourvecs) + radiative_return.py:18-24 (boost)."""
    me = m_electron
    s = 2.0 * me * (me + E)
    beta = (2.0 * alpha_em / math.pi) * (math.log(s / me ** 2) - 1.0)
    umax = (1.0 - mV ** 2 / s) ** (beta / 2.0)
    x1 = 1.0 - (x[0] * umax) ** (2.0 / beta)
    x2 = mV ** 2 / (x1 * s)
    rs = math.sqrt(s)
    E1, E2 = x1 * rs / 2.0, x2 * rs / 2.0
    pV = [E1 + E2, 0.0, 0.0, E1 - E2]
    p = [rs / 2.0, 0.0, 0.0, -math.sqrt(s / 4.0 - me ** 2)]
    # boost(p, v)
    lor = lambda a, b: a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3]
    rsq = math.sqrt(lor(p, p))
    v0 = lor(p, pV) / rsq
    c1 = (pV[0] + v0) / (rsq + p[0])
    out = [v0, pV[1] - c1 * p[1], pV[2] - c1 * p[2], pV[3] - c1 * p[3]]
    return [out, out]


def kin_compton_bound(E, x, mV, Pe, cte, u_az1, u_az2):
    """kinematics.py:134-183 (compton_fourvecs_boundelectron); returns [electron, V]."""
    me = m_electron
    Eg, ct = E, x[0]
    s = me ** 2 + 2 * Eg * (math.sqrt(me ** 2 + Pe ** 2) - cte * Pe)
    rs = math.sqrt(s)
    Ee = (s - mV ** 2 + me ** 2) / (2 * rs)
    EV = (s + mV ** 2 - me ** 2) / (2 * rs)
    pF = _sqrt(Ee ** 2 - me ** 2)
    bnum = math.sqrt(Eg ** 2 + 2 * cte * Eg * Pe + Pe ** 2)
    bden = Eg + math.sqrt(me ** 2 + Pe ** 2)
    b0 = bnum / bden
    g0 = 1.0 / math.sqrt(1.0 - b0 ** 2)
    ph = u_az1 * TWO_PI
    st = _sqrt(1 - ct ** 2)
    sp, cp = math.sin(ph), math.cos(ph)
    EeLab = g0 * Ee + b0 * g0 * pF * ct
    pe3 = [-pF * st * sp, -pF * st * cp, b0 * g0 * Ee + g0 * pF * ct]
    EVLab = g0 * EV - b0 * g0 * pF * ct
    pV3 = [pF * st * sp, pF * st * cp, b0 * g0 * EV - g0 * pF * ct]
    ctz = (Eg + cte * Pe) / math.sqrt(Eg ** 2 + 2 * cte * Eg * Pe + Pe ** 2)
    stz = _sqrt(1.0 - ctz ** 2)
    phie = u_az2 * TWO_PI
    ce, se = math.cos(phie), math.sin(phie)
    R = [[ctz * ce, -se, stz * ce], [ctz * se, ce, stz * se], [-stz, 0.0, ctz]]
    return [[EeLab] + rotate(R, pe3), [EVLab] + rotate(R, pV3)]


def boost_matrix(pf, mass):
    """particle.py:187-207."""
    E0, px, py, pz = pf
    gamma = E0 / mass
    beta = 1.0 if gamma == 1.0 else math.sqrt(1.0 - 1.0 / gamma ** 2)
    pmag = norm3((px, py, pz))
    if pmag == 0.0:
        return [[1.0 if i == j else 0.0 for j in range(4)] for i in range(4)]
    bx, by, bz = beta * px / pmag, beta * py / pmag, beta * pz / pmag
    g1 = gamma - 1
    b2 = beta ** 2
    return [[gamma, gamma * bx, gamma * by, gamma * bz],
            [gamma * bx, 1 + g1 * bx ** 2 / b2, g1 * bx * by / b2, g1 * bx * bz / b2],
            [gamma * by, g1 * by * bx / b2, 1 + g1 * by ** 2 / b2, g1 * by * bz / b2],
            [gamma * bz, g1 * bz * bx / b2, g1 * bz * by / b2, 1 + g1 * bz ** 2 / b2]]


def two_body_decay(pf, mX, m1, m2, u_cos, u_phi):
    """particle.py:209-256, isotropic: cos = U(-1,1), phi = U(0,2pi); returns two lab four-vectors."""
    E1 = (mX ** 2 - m2 ** 2 + m1 ** 2) / (2 * mX)
    E2 = (mX ** 2 - m1 ** 2 + m2 ** 2) / (2 * mX)
    pF = math.sqrt(E1 ** 2 - m1 ** 2)
    c = -1.0 + (1.0 - -1.0) * u_cos
    phi = 0.0 + (TWO_PI - 0.0) * u_phi
    sth = math.sqrt(1 - c ** 2)
    p1 = [E1, -pF * sth * math.sin(phi), -pF * sth * math.cos(phi), -pF * c]
    p2 = [E2, pF * sth * math.sin(phi), pF * sth * math.cos(phi), pF * c]
    B = boost_matrix(pf, mX)
    mv = lambda v: [sum(B[i][j] * v[j] for j in range(4)) for i in range(4)]
    return mv(p1), mv(p2)


def invariant_mass_rounded(p):
    """particle.py:125-131 (Q-7, Q-21)."""
    return round(float(np.sqrt(round(p[0] ** 2 - p[1] ** 2 - p[2] ** 2 - p[3] ** 2, 12))), 6)
