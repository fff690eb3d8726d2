"""SM shower stepping (oracle).  TEST INFRASTRUCTURE - see oracle/__init__.py.

Restates reference ``src/PETITE/shower.py``: table set-up :202-320, get_mfp :370-389, draw_sample :401-465,
sample_scattering :467-507, propagate_particle :509-601, generate_shower :603-708, and
``particle.py:363-424`` (short-lived two-body decays; decay in flight of the long-lived pi+-, K+-).  Reads this repo's ``data/*.npz`` tables (the same
numbers as the reference's pickles, repacked by tools/pack_reference_data.py) and the regenerated max_F.
"""
import math
import os

import numpy as np

from . import consts as C
from . import physics as phy
from .draws import CounterDraws, StreamDraws
from .findmax import split_grid
from .integrands import DSIGMA, sigma_moller, sigma_bhabha, sigma_muone, muone_threshold
from .philox import root_key, MCS_FINAL_INDEX
from .vegasmap import map_points

DATA_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "data", "")


class LinTable:
    """scipy ``interp1d(x, y, fill_value=f, bounds_error=False)`` (linear), scalar evaluation.

    Same arithmetic as scipy's _call_linear: hi = clip(searchsorted(x, xn), 1, n-1), lo = hi-1,
    y = (y_hi - y_lo)/(x_hi - x_lo) * (xn - x_lo) + y_lo; outside [x0, x_last] -> fill.
    """

    def __init__(self, x, y, fill=0.0):
        self.x = np.ascontiguousarray(x, dtype=np.float64)
        self.y = np.ascontiguousarray(y, dtype=np.float64)
        self.fill = fill

    def __call__(self, xn):
        x = self.x
        if not (xn >= x[0] and xn <= x[-1]):
            return self.fill if xn == xn else float("nan")
        hi = int(np.searchsorted(x, xn))
        hi = min(max(hi, 1), len(x) - 1)
        lo = hi - 1
        slope = (self.y[hi] - self.y[lo]) / (x[hi] - x[lo])
        return float(slope * (xn - x[lo]) + self.y[lo])


class OParticle:
    __slots__ = ("p0", "r0", "pf", "rf", "ended", "PID", "ID", "parent_PID", "parent_ID", "gen", "process",
                 "weight", "mass", "stability", "draws", "ntrials", "nsub", "parent_index")

    def __init__(self, p0, r0=(0.0, 0.0, 0.0), PID=11, ID=1, parent_PID=22, parent_ID=-1, gen=0, process="Input",
                 weight=1.0, mass=None, stability="stable"):
        self.p0 = [float(v) for v in p0]
        self.r0 = [float(v) for v in r0]
        self.pf = list(self.p0)
        self.rf = list(self.r0)
        self.ended = False
        self.PID, self.ID, self.parent_PID, self.parent_ID = PID, ID, parent_PID, parent_ID
        self.gen, self.process, self.weight, self.stability = gen, process, weight, stability
        self.mass = phy.invariant_mass_rounded(self.p0) if mass is None else mass   # particle.py:125-131 (Q-7)
        self.draws = None
        self.ntrials = 0
        self.nsub = 0
        self.parent_index = -1


class OracleShower:
    def __init__(self, dict_dir=None, target_material="graphite", min_energy=0.010, maxF_fudge_global=1,
                 max_n_integrators=int(1e4), seed=0, rescale_MCS=1, rng="counter"):
        self.dict_dir = DATA_DIR if dict_dir is None else dict_dir
        self.material = target_material
        self.min_energy = min_energy
        self.fudge = maxF_fudge_global
        self.max_sweeps = max_n_integrators
        self.seed = seed
        self.rescale_MCS = rescale_MCS
        self.rng = rng
        t = C.TARGETS[target_material]
        self.Z, self.A, self.rho, self.dEdx = t["Z_T"], t["A_T"], t["rho"], t["dEdx"]
        self.nT = self.rho / C.m_proton_grams / self.A       # shower.py:207
        self.ne = self.nT * self.Z                           # shower.py:208
        self._load_tables()

    # ---- shower.py:210-320 ----
    def _load_tables(self):
        xs = np.load(self.dict_dir + "sm_xsec.npz")
        mz = np.load(self.dict_dir + "sm_maps.npz")
        mf = np.load(self.dict_dir + "sm_maxF.npz")
        X = {P: xs[f"{P}/{self.material}"] for P in C.SM_PROCESSES}
        mue = X["MuonE"]
        while mue[0][1] == 0.0:                               # shower.py:238-239
            mue = mue[1:]
        X["MuonE"] = mue
        self.xsec = X
        self.min_calc = {                                     # shower.py:242-246
            11: min(X["Brem"][0][0], X["Moller"][0][0]),
            -11: min(X["Brem"][0][0], X["Bhabha"][0][0], X["Ann"][0][0]),
            13: max(min(X["MuonBrem"][0][0], X["MuonE"][0][0]), 0.120),
            -13: max(min(X["MuonBrem"][0][0], X["MuonE"][0][0]), 0.120),
            22: min(X["PairProd"][0][0], X["Comp"][0][0])}
        self.maps = {}
        for P in C.SM_PROCESSES:
            ninc = mz[f"{P}/ninc"]
            self.maps[P] = dict(E=mz[f"{P}/E"], ninc=ninc, grid=mz[f"{P}/grid"], meta=mz[f"{P}/meta"],
                                maxF=mf[f"{P}/{self.material}"])
        self.Eg_min = float(self.maps["Brem"]["meta"][1])     # shower.py:214-215
        self.Ee_min = float(self.maps["Brem"]["meta"][2])
        nZ, ne, G = self.nT, self.ne, C.GeVsqcm2
        NS = {}
        for P, n in (("Brem", nZ), ("PairProd", nZ), ("Ann", ne), ("Comp", ne), ("MuonBrem", nZ)):
            NS[P] = LinTable(X[P][:, 0], n * G * X[P][:, 1])  # shower.py:280-283,295
        BS = X["Brem"]
        bm_E = np.geomspace(3.0 * C.m_electron + self.Ee_min, BS[-1][0], len(BS))      # shower.py:285
        NS["Moller"] = LinTable(bm_E, ne * G * sigma_moller(bm_E, self.Ee_min))
        NS["Bhabha"] = LinTable(bm_E, ne * G * sigma_bhabha(bm_E, self.Ee_min))
        mu_min = float(muone_threshold(self.Ee_min))          # shower.py:289-291
        if mu_min < X["MuonE"][0][0]:
            mu_min = X["MuonE"][0][0]
        mu_E = np.geomspace(mu_min, bm_E[-1], len(BS))
        NS["MuonE"] = LinTable(mu_E, ne * G * sigma_muone(mu_E, self.Ee_min))
        self.NSigma = NS

    # ---- shower.py:357-389 ----
    def nsigma_total(self, PID, E):
        NS = self.NSigma
        if PID == 22:
            return NS["PairProd"](E) + NS["Comp"](E)
        if PID == 11:
            return NS["Brem"](E) + NS["Moller"](E)
        if PID == -11:
            return NS["Brem"](E) + NS["Bhabha"](E) + NS["Ann"](E)
        if abs(PID) == 13:
            return NS["MuonBrem"](E) + NS["MuonE"](E)
        raise ValueError(PID)

    def get_mfp(self, PID, E):
        ns = self.nsigma_total(PID, E)
        if ns <= 0.0:
            return 1.0e12
        return C.cmtom / ns

    # ---- shower.py:401-465 (VEGAS accept/reject; one-hypercube sweeps of B points, oracle/vegasmap.py) ----
    def lookup_key(self, mp, Einc):
        E = mp["E"]
        lu = int(np.argmin(np.abs(E - Einc))) + 1             # Q-1
        if lu >= len(E):
            lu = len(E) - 1
        return lu

    def event_info(self, process, Einc):
        ev = dict(E_inc=Einc, Z_T=self.Z, A_T=self.A, mT=self.A, mV=0.0, Eg_min=self.Eg_min, Ee_min=self.Ee_min)
        if process == "Brem":
            ev["m_lepton"] = C.m_electron
        if process == "MuonBrem":
            ev["m_lepton"] = C.m_muon
        return ev

    def _accept_reject(self, mp, lu, process, ev, draws, pc):
        grid = split_grid(mp["grid"][lu], mp["ninc"])
        dim = len(grid)
        B = int(mp["meta"][0])
        max_F = mp["maxF"][lu] * self.fudge
        f = DSIGMA[process]
        ntr = 0
        for sweep in range(self.max_sweeps):
            y = draws.vegas_y(sweep, B, dim, pc)
            x, jac = map_points(grid, y)
            wf = (jac / B) * f(x, ev)
            us = draws.vegas_u(sweep, B, dim, pc)
            if us is not None:
                hit = np.nonzero(max_F * us < wf)[0]
                if len(hit):
                    j = int(hit[0])
                    return x[j], ntr + j + 1
                ntr += B
            else:
                for j in range(B):
                    ntr += 1
                    if max_F * draws.accept_u() < wf[j]:
                        return x[j], ntr
        return None, ntr

    def draw_sample(self, Einc, process, draws, LU_Key=-1):
        mp = self.maps[process]
        lu = LU_Key
        if lu < 0 or lu > len(mp["E"]):
            lu = self.lookup_key(mp, Einc)
        x, ntr = self._accept_reject(mp, lu, process, self.event_info(process, Einc), draws, C.PROC_CODE[process])
        if x is None:
            raise Exception("No Sample Found", process, Einc, lu)
        return x, ntr

    # ---- shower.py:467-507 ----
    def sample_scattering(self, p, process):
        E0 = p.pf[0]
        if E0 <= max(self.min_calc[p.PID], self.min_energy, p.mass):
            return None
        RM = phy.rotation_matrix(p.pf)
        x, ntr = self.draw_sample(E0, process, p.draws)
        p.ntrials = ntr
        u = p.draws.kin(C.PROC_CODE[process])
        if process in ("Brem", "MuonBrem"):
            v1, v2 = phy.kin_brem(E0, p.mass, x, u[0])
        elif process == "PairProd":
            v1, v2 = phy.kin_pairprod(E0, x, u[0])
        elif process == "Comp":
            v1, v2 = phy.kin_compton(E0, x, u[0])
        elif process == "Ann":
            v1, v2 = phy.kin_annihilation(E0, x, u[0])
        elif process in ("Moller", "Bhabha"):
            v1, v2 = phy.kin_ee(E0, x, u[0])
        elif process == "MuonE":
            v1, v2 = phy.kin_mue(E0, x, u[0])
        else:
            raise ValueError(process)
        out = []
        for bit, (v, pid) in enumerate(zip((v1, v2), C.PROC_PIDS[process])):
            pid = p.PID if pid == 0 else pid
            lab = [v[0]] + phy.rotate(RM, v[1:])
            d = OParticle(lab, p.rf, PID=pid, ID=2 * p.ID + bit, parent_PID=p.PID, parent_ID=p.ID, gen=p.gen + 1,
                          process=process, weight=p.weight, mass=C.MASS[pid])
            d.draws = p.draws.child(bit)
            out.append(d)
        return out

    def _mcs(self, p4, dist_m, m_lepton, draws, index, pc=0):
        """moliere.py:350-400 through shower.py:545-547,579-581,596-598; no draws for a particle at rest."""
        if not phy.norm3(p4[1:]) > 0:
            return list(p4)
        s, z1, z2, up = draws.mcs(index, pc)
        return phy.mcs_scatter(p4, self.rho * (dist_m / C.cmtom), self.A, self.Z, self.rescale_MCS, m_lepton, s, z1, z2, up)

    # ---- shower.py:509-601 ----
    def propagate(self, p, losses, MS):
        if p.ended:
            return
        pmin = max(self.min_calc[p.PID], self.min_energy, p.mass)
        if p.p0[0] < pmin:
            p.ended = True
            return
        A, Z, rho = self.A, self.Z, self.rho
        d = p.draws
        if not losses:
            mfp = self.get_mfp(p.PID, p.pf[0])
            distC = d.final()
            dist = mfp * math.log(1.0 / (1.0 - distC))
            p3 = p.p0[1:]
            if MS:
                P0 = self._mcs(p.p0, dist, C.MASS[p.PID], d, MCS_FINAL_INDEX)
                q = [p3[0] + P0[1], p3[1] + P0[2], p3[2] + P0[3]]
                n = phy.norm3(q)
                hat = [q[0] / n, q[1] / n, q[2] / n]
                p.pf = P0
            else:
                n = phy.norm3(p3)
                hat = [p3[0] / n, p3[1] / n, p3[2] / n]
            p.rf = [p.r0[0] + hat[0] * dist, p.r0[1] + hat[1] * dist, p.r0[2] + hat[2] * dist]
        else:
            delta_z = 0
            hard = False
            i = 0
            m_l = C.MASS[p.PID]
            while not hard and p.pf[0] >= pmin:
                mfp = self.get_mfp(p.PID, p.pf[0])
                u_hard, u_dz = d.substep(i)
                delta_z = mfp / (6 + (20 - 6) * u_dz)
                if u_hard > math.exp(-delta_z / mfp):
                    hard = True
                else:
                    p.pf = phy.lose_energy(p.pf, p.mass, losses * delta_z)
                    pn = phy.norm3(p.pf[1:])
                    if pn > 0.0:
                        p.rf = [p.rf[k] + p.pf[1 + k] / pn * delta_z for k in range(3)]
                    if MS:
                        p.pf = self._mcs(p.pf, delta_z, m_l, d, i)
                    p.nsub += 1
                i += 1
            distC = d.final()
            if p.pf[0] < pmin:
                last = distC * delta_z
            else:
                mfp = self.get_mfp(p.PID, p.pf[0])
                last = mfp * math.log(1.0 / (1.0 + (math.exp(-delta_z / mfp) - 1) * distC))
            p.pf = phy.lose_energy(p.pf, p.mass, losses * last)
            pn = phy.norm3(p.pf[1:])
            if pn > 0.0:
                p.rf = [p.rf[k] + p.pf[1 + k] / pn * last for k in range(3)]
            if MS:
                # Q-12: electron mass regardless of the lepton
                p.pf = self._mcs(p.pf, last, C.m_electron, d, MCS_FINAL_INDEX)
        p.ended = True

    # ---- particle.py:363-389 (draw_x_sample, prob_decay_b_int) ----
    @staticmethod
    def _decay_rates(p, int_length, ctau0):
        E = p.p0[0]                                      # particle.py:358, 381: the energy at creation
        gamma = E / p.mass
        beta = np.sqrt(1 - 1 / gamma ** 2)
        ctau = ctau0 * gamma * beta
        return ctau, 1.0 / ctau + 1.0 / int_length

    def _draw_x(self, p, loop, tot_rate):
        c = tot_rate                                     # decay_int_prob(0)[0]
        x_max = 4 / (tot_rate * np.exp(-tot_rate * 0))   # 4 / decay_int_prob(0)[1]
        i = 0
        while True:
            u1, u2 = p.draws.decay_x(loop, i)
            x = 0 + (x_max - 0) * u1                     # np.random.uniform(0, x_max)
            if u2 < tot_rate * np.exp(-tot_rate * x) / c:
                return x
            i += 1

    # ---- particle.py:391-424 ----
    def decay(self, p):
        if p.PID not in C.MESON_DECAYS:
            raise ValueError("Decay options for particle not specified.")
        opts = C.MESON_DECAYS[p.PID]
        if len(opts) != 1:
            raise NotImplementedError("multi-channel decays are outside the hot-path scope (SURVEY.md 2)")
        br, dec = opts[0]
        if len(dec) != 2 or p.stability not in ("short-lived", "long-lived"):
            raise ValueError("only two-body decays are in scope")
        weights = [p.weight * br, p.weight * br]
        if p.stability == "long-lived":                  # particle.py:410-422: decay in flight of pi+-, K+-
            int_length, ctau0 = C.INT_LENGTH[p.PID], C.DECAY_LENGTH[p.PID]
            ctau, tot_rate = self._decay_rates(p, int_length, ctau0)
            delta_z = self._draw_x(p, 1, tot_rate)
            pf0 = float(np.linalg.norm(p.pf[1:]))
            if pf0 > 0.0:
                p.rf = [p.rf[k] + p.pf[1 + k] / pf0 * delta_z for k in range(3)]
            for bit in range(2):                         # prob_decay_b_int draws its own x for EACH daughter dictionary
                x = self._draw_x(p, 2 + bit, tot_rate)
                weights[bit] = p.weight * br * ((1 - np.exp(-tot_rate * x)) / (1 + ctau / int_length))
        uc, up = p.draws.decay()
        v1, v2 = phy.two_body_decay(p.pf, p.mass, C.MASS[dec[0]], C.MASS[dec[1]], uc, up)
        out = []
        for bit, (v, pid) in enumerate(zip((v1, v2), dec)):
            dd = OParticle(v, p.rf, PID=pid, ID=2 * p.ID + bit, gen=p.gen + 1, process="SMDecay",
                           weight=weights[bit], mass=C.MASS[pid])      # Q-10: parent_PID/parent_ID defaults
            dd.draws = p.draws.child(bit)
            out.append(dd)
        p.ended = True
        return out

    # ---- shower.py:665-698 ----
    CHOICES = {11: ["Brem", "Moller"], -11: ["Brem", "Ann", "Bhabha"], 22: ["PairProd", "Comp"],
               13: ["MuonE", "MuonBrem"], -13: ["MuonE", "MuonBrem"]}

    def choose_process(self, p):
        labels = self.CHOICES[p.PID]
        c = np.array([self.NSigma[l](p.pf[0]) for l in labels])
        SC = np.sum(c)
        if SC == 0.0 or np.isnan(SC):
            return None
        pr = c / SC
        u = p.draws.choice()
        cdf = pr.cumsum()
        cdf /= cdf[-1]
        return labels[int(cdf.searchsorted(u, side="right"))]

    # ---- shower.py:603-708 ----
    def generate_shower(self, p0, shower_id=0, GlobalMS=True):
        root = OParticle(p0.p0, p0.r0, PID=p0.PID, ID=p0.ID, parent_PID=p0.parent_PID, parent_ID=p0.parent_ID,
                         gen=p0.gen, process=p0.process, weight=p0.weight, mass=p0.mass, stability=p0.stability)
        root.draws = CounterDraws(root_key(self.seed, shower_id)) if self.rng == "counter" else StreamDraws()
        allp = [root]
        MS_e = bool(GlobalMS)
        if root.p0[0] < self.min_energy:
            return allp
        dEdxT = self.dEdx * 0.1
        i = 0
        n_open = 1
        while i < len(allp):
            ap = allp[i]
            idx = i
            i += 1
            if ap.ended:
                continue
            new = None
            if ap.stability in ("short-lived", "long-lived"):
                new = self.decay(ap)
                n_open -= 1
            elif ap.stability == "stable":
                if ap.PID == 22:
                    self.propagate(ap, False, False)
                elif abs(ap.PID) in (11, 13):
                    self.propagate(ap, dEdxT, MS_e)
                if ap.ended:
                    n_open -= 1
                if n_open == 0 and ap.pf[0] < self.min_energy:
                    break
                if ap.PID in self.CHOICES:
                    proc = self.choose_process(ap)
                    if proc is None:
                        continue
                    new = self.sample_scattering(ap, proc)
                elif abs(ap.PID) == 14:
                    ap.ended = True
                    n_open -= 1
                else:
                    raise ValueError("Q-20: stable PID %d would loop forever in the reference" % ap.PID)
            if new is None:
                continue
            for dp in new:
                if dp.p0[0] > self.min_energy:
                    dp.parent_index = idx
                    allp.append(dp)
                    n_open += 1
        return allp
