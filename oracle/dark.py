"""Dark-vector pass over an SM shower (oracle).  TEST INFRASTRUCTURE - see oracle/__init__.py.

Restates reference ``src/PETITE/dark_shower.py``: GetBSMWeights :595-647, draw_dark_sample :649-704,
draw_pe_sample :706-719, produce_bsm_particle :721-804, generate_dark_shower :806-849, for the default
``bound_electron=True`` configuration.  The constructor's set-up tables (weights, dRate/dE, n*sigma_dark; nested
quadratures at dark_shower.py:254-593) are NOT recomputed here: they are read from ``data/dark_setup_*.npz``, which
tests/golden/make_golden.py dumps from the UNMODIFIED reference constructor.
"""
import math

import numpy as np

from . import consts as C
from . import physics as phy
from .findmax import split_grid
from .integrands import DSIGMA
from .shower import OracleShower, OParticle, LinTable, DATA_DIR
from .philox import MCS_FINAL_INDEX
from .vegasmap import map_points

PID_V = 4900022
ALL_DARK = ["DarkBrem", "DarkAnn", "DarkComp", "TwoBody_BSMDecay", "DarkMuonBrem"]


def mv_tag(mV):
    return repr(float(mV))


class LogLogTable:
    """dark_shower.py:31-46 (interpolate1d, xspace = yspace = 'log', fill_value = -20): tables hold log10 values."""

    def __init__(self, lx, ly):
        self.t = LinTable(lx, ly, fill=-20.0)

    def __call__(self, x):
        return 10 ** self.t(math.log10(x))


class OracleDarkShower(OracleShower):
    def __init__(self, dict_dir=None, target_material="graphite", min_energy=0.010, mV_in_GeV=0.03, active_processes=None,
                 **kw):
        super().__init__(dict_dir, target_material, min_energy, **kw)
        self.active = ALL_DARK if active_processes is None else list(active_processes)
        z = np.load((DATA_DIR if dict_dir is None else dict_dir) + f"dark_setup_{target_material}_mV{mv_tag(mV_in_GeV)}.npz")
        m = z["meta"]
        self.mV, self.mV_est, self.E_res, self.E_thr_comp, self.g_e, self.eps, self.Zeff = [float(v) for v in m[:7]]
        self.bound = bool(m[7])
        if not self.bound:
            raise NotImplementedError("bound_electron=False is outside the accelerated scope")
        self.min_dark = {(int(a), str(b)): float(c) for a, b, c in zip(z["min_dark_pid"], z["min_dark_proc"], z["min_dark_E"])}
        self.weights = {k: LinTable(z[f"weights/{k}"][:, 0], z[f"weights/{k}"][:, 1]) for k in
                        ("brem_elec", "brem_positron", "muon_brem", "annihilation")}
        self.drate = {k: (z[f"drate/{k}/E"], z[f"drate/{k}/table"]) for k in ("brem_elec", "brem_positron", "muon_brem", "annihilation")}
        self.NSdark = {P: LogLogTable(z[f"nsdark/{P}/x"], z[f"nsdark/{P}/y"]) for P in C.DARK_PROCESSES}
        dm = np.load(self.dict_dir + f"dark_maps_mV{mv_tag(self.mV_est)}.npz")
        mf = np.load(self.dict_dir + "dark_maxF.npz")
        self.dmaps = {P: dict(E=dm[f"{P}/E"], ninc=dm[f"{P}/ninc"], grid=dm[f"{P}/grid"], meta=dm[f"{P}/meta"],
                              maxF=mf[f"{mv_tag(self.mV_est)}/{P}/{target_material}"]) for P in C.DARK_PROCESSES}

    # ---- dark_shower.py:595-647 ----
    def bsm_weight(self, PID, E0, mass, process):
        if PID not in (-11, 11, 13, -13, 22, 111, 221, 331):
            return 0.0
        if (PID, process) not in self.min_dark:
            return 0.0
        if E0 < self.min_dark[(PID, process)]:
            return 0.0
        pre = self.g_e ** 2 / (4 * math.pi * C.alpha_em)
        if PID == 22:
            if process != "DarkComp" or E0 < self.min_calc[22]:
                return 0.0
            with np.errstate(all="ignore"):       # numpy semantics: x/0 -> inf, as in the reference
                return float(np.float64(pre * self.NSdark["DarkComp"](E0)) / np.float64(self.NSigma["PairProd"](E0) + self.NSigma["Comp"](E0)))
        if process == "DarkBrem":
            if abs(PID) != 11:
                return 0.0
            return pre * self.weights["brem_elec" if PID == 11 else "brem_positron"](E0)
        if PID == -11 and process == "DarkAnn":
            return pre * self.weights["annihilation"](E0)
        if PID in (111, 221, 331):
            if process == "TwoBody_BSMDecay":
                r = self.mV / mass
                if r >= 1.0:
                    return 0.0
                return 2 * self.eps ** 2 * (1.0 - r ** 2) ** 3 * C.MESON_DECAYS[PID][0][0]
            return 0.0
        if abs(PID) == 13 and process == "DarkMuonBrem":
            return pre * self.weights["muon_brem"](E0)
        return 0.0

    # ---- dark_shower.py:649-704 ----
    def draw_dark_sample(self, Einc, process, draws):
        mp = self.dmaps[process]
        lu = self.lookup_key(mp, Einc)
        ev = dict(E_inc=Einc, Z_T=self.Z, A_T=self.A, mT=self.A, mV=self.mV, Eg_min=self.Eg_min,
                  m_lepton=C.m_muon if process == "DarkMuonBrem" else C.m_electron)
        return self._accept_reject(mp, lu, process, ev, draws, C.PROC_CODE[process])

    def _wave_function(self, pe):
        lam = C.alpha_em * self.Zeff * C.m_electron
        return 32 / math.pi * lam ** 5 * pe ** 2 / (pe ** 2 + lam ** 2) ** 4

    def draw_pe(self, draws, pc):
        c = self._wave_function(C.alpha_em * self.Zeff * C.m_electron / math.sqrt(3))
        i = 0
        while True:
            ux, uu = draws.pe(i, pc)
            x = 0 + (1e-3 - 0) * ux
            i += 1
            if uu < self._wave_function(x) / c:
                return x

    # ---- dark_shower.py:721-804 ----
    def produce_bsm_particle(self, p, process, wg):
        pc = C.PROC_CODE[process]
        d = p.draws
        pf, mass = list(p.pf), p.mass
        name = None
        if process == "DarkAnn" and p.PID == -11:
            name = "annihilation"
        elif process == "DarkBrem":
            name = "brem_elec" if p.PID == 11 else "brem_positron"
        elif process == "DarkMuonBrem":
            name = "muon_brem"
        if name is not None:
            Es, tab = self.drate[name]
            E0 = p.p0[0]
            if E0 < Es.min():
                ie = int(np.argmin(Es))
            elif E0 > Es.max():
                ie = int(np.argmax(Es))
            else:
                ok = np.nonzero(Es <= E0)[0]
                ie = int(ok[np.argmax(Es[ok])])
            Ei = Es[ie]
            energies, rel = tab[ie][:, 0], tab[ie][:, 1]
            if np.sum(rel) == 0.0:
                return None
            rel = rel / np.sum(rel)
            cdf = rel.cumsum()
            cdf /= cdf[-1]
            E_int = energies[int(cdf.searchsorted(d.dbin(pc), side="right"))] + (E0 - Ei)
            dist = (p.p0[0] - E_int) / (self.dEdx * 0.1)
            pf = self._mcs(p.p0, dist, C.m_electron, d, MCS_FINAL_INDEX, pc)      # Q-12: default m_lepton
            pf = phy.lose_energy(pf, mass, E0 - E_int)
        E0 = pf[0]
        RM = phy.rotation_matrix(pf)
        ntr = 0
        if process == "DarkAnn" and E0 <= self.E_res:
            V = [math.sqrt(E0 ** 2 - C.m_electron ** 2 + self.mV ** 2), 0, 0, math.sqrt(E0 ** 2 - C.m_electron ** 2)]
        elif process == "DarkComp" and E0 <= self.E_thr_comp:
            V = [math.sqrt(E0 ** 2 + self.mV ** 2), 0, 0, E0]
        else:
            x, ntr = self.draw_dark_sample(E0, process, d)
            if x is None:
                return None
            if process == "DarkComp":
                pe = self.draw_pe(d, pc)
                c0 = -1 + (1 - -1) * d.c0(pc)
                s = C.m_electron ** 2 + 2 * E0 * (math.sqrt(C.m_electron ** 2 + pe ** 2) - c0 * pe)
                Ee = (s - self.mV ** 2 + C.m_electron ** 2) / (2 * math.sqrt(s))
                if Ee < C.m_electron:
                    V = [self.mV, 0, 0, 0]
                    wg = 0.0
                else:
                    u = d.kin(pc)
                    V = phy.kin_compton_bound(E0, x, self.mV, pe, c0, u[0], u[1])[-1]
            elif process == "DarkAnn":
                V = phy.kin_darkann(E0, x, self.mV)[-1]
            else:
                u = d.kin(pc)
                V = phy.kin_darkbrem(E0, mass, x, u[0], self.mV)[-1]
        lab = [V[0]] + phy.rotate(RM, V[1:])
        gp = {"DarkAnn": "DarkAnn_bound", "DarkComp": "DarkComp_bound"}.get(process, process)
        v = OParticle(lab, p.rf, PID=PID_V, ID=2 * p.ID, parent_PID=p.PID, parent_ID=p.ID, gen=p.gen + 1, process=gp,
                      weight=wg * p.weight, mass=None)
        v.ntrials = ntr
        return v

    # ---- dark_shower.py:806-849 ----
    def generate_dark_shower(self, sm_shower):
        out = []
        for idx, ap in enumerate(sm_shower):
            for process in self.active:
                wg = self.bsm_weight(ap.PID, ap.p0[0], ap.mass, process)
                if wg > 0.0:
                    if process == "TwoBody_BSMDecay":
                        uc, up = ap.draws.decay(C.PROC_CODE["TwoBody_BSMDecay"])
                        _, v4 = phy.two_body_decay(ap.pf, ap.mass, 0, self.mV, uc, up)
                        v = OParticle(v4, ap.rf, PID=PID_V, ID=2 * ap.ID + 1, parent_PID=ap.PID, parent_ID=ap.ID,
                                      gen=ap.gen + 1, process=process, weight=ap.weight * wg, mass=self.mV)
                    else:
                        v = self.produce_bsm_particle(ap, process, wg)
                    if v is not None:
                        v.parent_index = idx
                        out.append(v)
        return sm_shower, out
