#!/usr/bin/env python
"""Generate the golden vectors under tests/golden/ by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference; the GPU box does not have it):

    python tests/golden/make_golden.py

The reference is imported through ``_refstub`` (stub ``vegas``/``matplotlib`` modules, see its docstring).
Everything recorded here is computed by the reference's own functions:

    integrands.npz   dsigma_* classes of all_processes.py at points drawn through the shipped maps
    kinematics.npz   *_fourvecs of kinematics.py (azimuth uniform recorded by re-seeding NumPy)
    mcs.npz          moliere.get_scattered_momentum_fast with its ``random`` calls fed from a tape
    particle.npz     Particle.lose_energy / rotation_matrix / two_body_decay
    nsigma.npz       Shower._NSigma*, get_mfp, thresholds for graphite and lead
    showers.npz      whole generate_shower runs (stream mode, seeded) - multiplicities, ids, four-vectors
    dark_kinematics.npz  l_to_lV_fourvecs / compton_fourvecs_boundelectron / radiative_return_fourvecs of kinematics.py
"""
import os
import pickle
import random
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, "..", ".."))
import _refstub  # noqa: E402

_refstub.import_reference()
warnings.filterwarnings("ignore")

import PETITE.all_processes as ap  # noqa: E402
import PETITE.kinematics as kin  # noqa: E402
import PETITE.moliere as mol  # noqa: E402
from PETITE.particle import Particle  # noqa: E402
from PETITE.physical_constants import m_electron, m_muon, alpha_em, m_pi0  # noqa: E402
from oracle.vegasmap import AdaptiveMapStub, map_points  # noqa: E402
from oracle.findmax import split_grid, MATERIALS  # noqa: E402
from oracle.consts import SM_PROCESSES, DARK_PROCESSES, TARGETS  # noqa: E402

DATA = os.path.join(HERE, "..", "..", "data", "")
REFDIR = "/tmp/petite_refdata/"


def build_reference_dict_dir(data=None, refdir=None, ref_sub="data", materials=None):
    """Reference-format dict_dir: shipped pickles + sm_maps.pkl / dark_maps.pkl rebuilt from <data>/*.npz.  Defaults: the 10 GeV
    set (data/ -> /tmp/petite_refdata/); ``data=data_400GeV/, ref_sub="data_400GeV"`` builds the 400 GeV one (max_F exists for the
    materials tools/make_400GeV.py was run for)."""
    data = DATA if data is None else data
    refdir = REFDIR if refdir is None else refdir
    materials = MATERIALS if materials is None else materials
    os.makedirs(refdir, exist_ok=True)
    import shutil
    for f in ["sm_xsec.pkl", "dark_xsec.pkl"]:
        if not os.path.exists(refdir + f):
            os.symlink(os.path.join(_refstub.REF_ROOT, ref_sub, f), refdir + f)
    for f in ["dark_weights.pkl", "dark_drate.pkl"]:      # the constructors write into these (SURVEY Q-3): real copies
        if not os.path.exists(refdir + f) or os.path.islink(refdir + f):
            if os.path.islink(refdir + f):
                os.unlink(refdir + f)
            shutil.copy(os.path.join(_refstub.REF_ROOT, ref_sub, f), refdir + f)
            os.chmod(refdir + f, 0o644)

    def mk(z, mf, procs, pre=""):
        out = {}
        for P in procs:
            E, ninc, G, meta = z[f"{P}/E"], z[f"{P}/ninc"], z[f"{P}/grid"], z[f"{P}/meta"]
            out[P] = [[float(E[i]), {"neval": int(meta[0]), "max_F": {m: float(mf[f"{pre}{P}/{m}"][i]) for m in materials},
                                     "adaptive_map": AdaptiveMapStub(split_grid(G[i], ninc)),
                                     "Eg_min": float(meta[1]), "Ee_min": float(meta[2])}] for i in range(len(E))]
        return out
    pickle.dump(mk(np.load(data + "sm_maps.npz"), np.load(data + "sm_maxF.npz"), SM_PROCESSES), open(refdir + "sm_maps.pkl", "wb"))
    dmf = np.load(data + "dark_maxF.npz")
    dm = {}
    for f in sorted(os.listdir(data)):
        if f.startswith("dark_maps_mV") and not f.endswith("_shipped.npz"):
            tag = f[len("dark_maps_mV"):-4]
            dm[float(tag)] = mk(np.load(data + f), dmf, DARK_PROCESSES, pre=tag + "/")
            # the reference's default active_processes includes "TwoBody_BSMDecay" and set_dark_samples
            # (dark_shower.py:196-201) looks every active process up in dark_maps.pkl: the lost pickle must have
            # carried such a key.  An empty entry lets the unmodified constructor run.
            dm[float(tag)]["TwoBody_BSMDecay"] = []
    pickle.dump(dm, open(refdir + "dark_maps.pkl", "wb"))


REF_DS = {"Brem": ap.dsigma_brem_dimensionless, "MuonBrem": ap.dsigma_brem_dimensionless,
          "PairProd": ap.dsigma_pairprod_dimensionless, "Comp": ap.dsigma_compton_dCT, "Ann": ap.dsigma_annihilation_dCT,
          "Moller": ap.dsigma_moller_dCT, "Bhabha": ap.dsigma_bhabha_dCT, "MuonE": ap.dsigma_muonelectron_dCT,
          "DarkBrem": ap.dsig_dx_dcostheta_dark_brem_exact_tree_level,
          "DarkMuonBrem": ap.dsig_dx_dcostheta_dark_brem_exact_tree_level,
          "DarkAnn": ap.dsigma_radiative_return_du, "DarkComp": ap.dsigma_compton_dCT}


def golden_integrands(rng):
    out = {}
    sm = np.load(DATA + "sm_maps.npz")
    dk = np.load(DATA + "dark_maps_mV0.03.npz")
    for material in ("graphite", "lead"):
        t = TARGETS[material]
        for P in SM_PROCESSES + DARK_PROCESSES:
            z = sm if P in SM_PROCESSES else dk
            E, ninc, G = z[f"{P}/E"], z[f"{P}/ninc"], z[f"{P}/grid"]
            rows = [3, 20, 45, 70, 99]
            Es, xs, fs = [], [], []
            for ie in rows:
                grid = split_grid(G[ie], ninc)
                y = rng.random((24, len(grid)))
                x, _ = map_points(grid, y)
                # sampler-time event_info (shower.py:435-439, dark_shower.py:672-676): mT = A_T
                for Einc in (float(E[ie]), float(E[ie]) * 0.93):
                    ev = {"E_inc": Einc, "m_e": m_electron, "Z_T": t["Z_T"], "A_T": t["A_T"], "mT": t["A_T"],
                          "alpha_FS": alpha_em, "mV": 0.0 if P in SM_PROCESSES else 0.03, "Eg_min": 0.001, "Ee_min": 0.005}
                    if P in ("Brem", "DarkBrem"):
                        ev["m_lepton"] = m_electron
                    if P in ("MuonBrem", "DarkMuonBrem"):
                        ev["m_lepton"] = m_muon
                    f = REF_DS[P](event_info=ev, ndim=len(grid))
                    for xi in x:
                        Es.append(Einc)
                        xs.append(np.pad(xi, (0, 4 - len(xi))))
                        fs.append(float(np.asarray(f(xi))))
            out[f"{material}/{P}/E"] = np.array(Es)
            out[f"{material}/{P}/x"] = np.array(xs)
            out[f"{material}/{P}/f"] = np.array(fs)
    np.savez_compressed(os.path.join(HERE, "integrands.npz"), **out)


def golden_kinematics(rng):
    out = {}
    sm = np.load(DATA + "sm_maps.npz")
    cases = [("Brem", kin.e_to_egamma_fourvecs, 11, m_electron), ("MuonBrem", kin.e_to_egamma_fourvecs, 13, m_muon),
             ("PairProd", kin.gamma_to_epem_fourvecs, 22, 0.0), ("Comp", kin.compton_fourvecs, 22, 0.0),
             ("Ann", kin.annihilation_fourvecs, -11, m_electron), ("Moller", kin.ee_to_ee_fourvecs, 11, m_electron),
             ("Bhabha", kin.ee_to_ee_fourvecs, -11, m_electron), ("MuonE", kin.mue_to_mue_fourvecs, 13, m_muon)]
    for P, fn, pid, mass in cases:
        E, ninc, G = sm[f"{P}/E"], sm[f"{P}/ninc"], sm[f"{P}/grid"]
        inp, res = [], []
        for ie in (25, 50, 75, 99):
            grid = split_grid(G[ie], ninc)
            x, _ = map_points(grid, rng.random((16, len(grid))))
            for xi in x:
                Einc = float(E[ie])
                p = Particle([Einc, 0, 0, np.sqrt(Einc ** 2 - mass ** 2)], [0, 0, 0], {"PID": pid, "mass": mass})
                seed = int(rng.integers(1 << 30))
                np.random.seed(seed)
                u = np.random.random()
                np.random.seed(seed)
                v = fn(p, xi)
                inp.append([Einc, mass] + list(np.pad(xi, (0, 4 - len(xi)))) + [u, 0.0])
                res.append(list(v[0]) + list(v[1]))
        out[f"{P}/in"] = np.array(inp, dtype=float)
        out[f"{P}/out"] = np.array(res, dtype=float)
    # pi0 -> gamma gamma (particle.py:209-256)
    inp, res = [], []
    for _ in range(64):
        pv = rng.normal(size=3) * rng.choice([0.05, 1.0, 30.0])
        E = np.sqrt(m_pi0 ** 2 + pv @ pv)
        p = Particle([E, *pv], [0.1, 0.2, 0.3], {"PID": 111, "mass": m_pi0, "stability": "short-lived"})
        seed = int(rng.integers(1 << 30))
        np.random.seed(seed)
        u1, u2 = np.random.random(), np.random.random()
        np.random.seed(seed)
        d = p.decay_particle()
        inp.append([E, m_pi0, pv[0], pv[1], pv[2], 0.0, u1, u2])
        res.append(list(d[0].get_p0()) + list(d[1].get_p0()))
    out["SMDecay/in"] = np.array(inp)
    out["SMDecay/out"] = np.array(res)
    np.savez_compressed(os.path.join(HERE, "kinematics.npz"), **out)


class _Tape:
    """Feeds moliere.py's ``random.choice / gauss / uniform`` from explicit values."""

    def __init__(self, sign, z1, z2, uphi):
        self.sign, self.z, self.uphi = sign, [z1, z2], uphi

    def choice(self, seq):
        return self.sign

    def gauss(self, mu, sigma):
        return mu + self.z.pop(0) * sigma

    def uniform(self, a, b):
        return a + (b - a) * self.uphi


def golden_mcs(rng):
    inp, res = [], []
    real_random = mol.random
    for material in ("graphite", "lead"):
        t = TARGETS[material]
        for _ in range(96):
            m = rng.choice([m_electron, m_muon])
            pmag = 10 ** rng.uniform(-2.5, 2)
            d = rng.normal(size=3)
            if rng.random() < 0.3:
                d = np.array([rng.normal() * 1e-4, rng.normal() * 1e-4, rng.choice([-1.0, 1.0])])
            if rng.random() < 0.1:
                d[rng.integers(3)] = 0.0
            d = d / np.linalg.norm(d) * pmag
            p4 = np.array([np.sqrt(pmag ** 2 + m ** 2), *d])
            dist = 10 ** rng.uniform(-5, -1)
            sign, z1, z2, uphi = float(rng.choice([-1, 1])), rng.normal(), rng.normal(), rng.random()
            mol.random = _Tape(sign, z1, z2, uphi)
            q = mol.get_scattered_momentum_fast(p4, t["rho"] * (dist / 0.01), t["A_T"], t["Z_T"], 1, m_lepton=m)
            inp.append(list(p4) + [dist, m, sign, z1, z2, uphi, t["Z_T"]])
            res.append(list(q))
    mol.random = real_random
    np.savez_compressed(os.path.join(HERE, "mcs.npz"), inp=np.array(inp), out=np.array(res))


def golden_particle(rng):
    le_in, le_out, rm_in, rm_out = [], [], [], []
    for _ in range(64):
        m = rng.choice([m_electron, m_muon])
        pv = rng.normal(size=3) * 10 ** rng.uniform(-3, 1)
        E = np.sqrt(m ** 2 + pv @ pv)
        p = Particle([E, *pv], [0, 0, 0], {"PID": 11, "mass": m})
        loss = E * rng.choice([1e-3, 0.3, 2.0])
        p.lose_energy(loss)
        le_in.append([E, *pv, m, loss])
        le_out.append(list(p.get_pf()))
        q = Particle([E, *pv], [0, 0, 0], {"PID": 11, "mass": m})
        rm_in.append([E, *pv])
        rm_out.append(np.array(q.rotation_matrix(), dtype=float).ravel())
    np.savez_compressed(os.path.join(HERE, "particle.npz"), lose_in=np.array(le_in), lose_out=np.array(le_out),
                        rot_in=np.array(rm_in), rot_out=np.array(rm_out))


def golden_nsigma():
    from PETITE.shower import Shower
    out = {}
    names = ["Brem", "Ann", "PairProd", "Comp", "Moller", "Bhabha", "MuonE", "MuonBrem"]
    attr = {"Brem": "_NSigmaBrem", "Ann": "_NSigmaAnn", "PairProd": "_NSigmaPP", "Comp": "_NSigmaComp",
            "Moller": "_NSigmaMoller", "Bhabha": "_NSigmaBhabha", "MuonE": "_NSigmaMuonE", "MuonBrem": "_NSigmaMuonBrem"}
    E = np.concatenate([np.geomspace(1e-3, 150.0, 400), [0.0016, 0.01, 1.0, 10.0, 100.0, 0.2, 0.25708651005470756]])
    for material in ("graphite", "lead"):
        s = Shower(REFDIR, material, 0.010)
        out[f"{material}/E"] = E
        for P in names:
            out[f"{material}/{P}"] = np.array([float(getattr(s, attr[P])(e)) for e in E])
            t = getattr(s, attr[P])
            out[f"{material}/{P}/table_x"], out[f"{material}/{P}/table_y"] = np.asarray(t.x), np.asarray(t.y)
        for pid in (22, 11, -11, 13):
            out[f"{material}/mfp/{pid}"] = np.array([float(s.get_mfp([pid, e])) for e in E])
        out[f"{material}/min_calc"] = np.array([s._minimum_calculable_energy[k] for k in (11, -11, 22, 13, -13)])
        out[f"{material}/n"] = np.array(s.get_n_targets())
    np.savez_compressed(os.path.join(HERE, "nsigma.npz"), **out)


def golden_showers():
    from PETITE.shower import Shower
    out = {}
    proc_code = {"Input": 15, "SMDecay": 12}
    proc_code.update({p: i for i, p in enumerate(SM_PROCESSES)})
    cases = [("graphite", 11, 10.0, 0.010, 101, None), ("graphite", 22, 10.0, 0.010, 102, None),
             ("lead", 22, 5.0, 0.010, 103, None), ("lead", -11, 3.0, 0.010, 104, None),
             ("graphite", 13, 20.0, 0.030, 105, m_muon), ("graphite", 111, 8.0, 0.010, 106, m_pi0),
             ("graphite", 11, 1.0, 0.010, 107, None), ("lead", 11, 0.5, 0.010, 108, m_electron)]
    for k, (mat, pid, E, Emin, seed, mass) in enumerate(cases):
        s = Shower(REFDIR, mat, Emin)
        m = {11: m_electron, -11: m_electron, 22: 0.0, 13: m_muon, -13: m_muon, 111: m_pi0}[pid]
        ids = {"PID": pid, "ID": 1}
        if mass is not None:
            ids["mass"] = mass
        if pid == 111:
            ids["stability"] = "short-lived"
        np.random.seed(seed)
        random.seed(seed)
        sh = s.generate_shower(Particle([E, 0, 0, np.sqrt(E ** 2 - m ** 2)], [0, 0, 0], ids))
        out[f"{k}/case"] = np.array([pid, E, Emin, seed, -1.0 if mass is None else mass])
        out[f"{k}/material"] = np.array(mat)
        out[f"{k}/pid"] = np.array([p.get_ids()["PID"] for p in sh])
        out[f"{k}/ID_mod"] = np.array([p.get_ids()["ID"] % (1 << 61) for p in sh], dtype=np.int64)
        out[f"{k}/gen"] = np.array([p.get_ids()["generation_number"] for p in sh])
        out[f"{k}/process"] = np.array([proc_code[p.get_ids()["generation_process"]] for p in sh])
        out[f"{k}/weight"] = np.array([p.get_ids()["weight"] for p in sh])
        out[f"{k}/mass"] = np.array([p.get_ids()["mass"] for p in sh], dtype=float)
        out[f"{k}/p0"] = np.array([p.get_p0() for p in sh], dtype=float)
        out[f"{k}/pf"] = np.array([p.get_pf() for p in sh], dtype=float)
        out[f"{k}/r0"] = np.array([p.get_r0() for p in sh], dtype=float)
        out[f"{k}/rf"] = np.array([p.get_rf() for p in sh], dtype=float)
        print("shower case", k, mat, pid, E, "->", len(sh), "particles")
    out["n_cases"] = np.array(len(cases))
    np.savez_compressed(os.path.join(HERE, "showers.npz"), **out)


DARK_CASES = [("graphite", 0.03), ("lead", 0.03), ("graphite", 0.003), ("graphite", 1.0)]


def dump_dark_setup(d, material, mV):
    """Constructor tables of a reference DarkShower -> data/dark_setup_<material>_mV<tag>.npz (set-up cache + oracle input)."""
    out = {"meta": np.array([d._mV, d._mV_estimator, d._resonant_annihilation_energy, d._compton_threshold_energy, d.g_e,
                             d.kinetic_mixing, d.Zeff, 1.0 if d.bound_electron else 0.0])}
    pids, procs, Es = [], [], []
    for pid, dd in d._minimum_calculable_dark_energy.items():
        for pr, e in dd.items():
            pids.append(pid); procs.append(pr); Es.append(float(e))
    out["min_dark_pid"], out["min_dark_proc"], out["min_dark_E"] = np.array(pids), np.array(procs), np.array(Es)
    for name, itp in (("brem_elec", d._brem_elec_numerical_weight), ("brem_positron", d._brem_positron_numerical_weight),
                      ("muon_brem", d._muon_brem_numerical_weight), ("annihilation", d._annihilation_numerical_weight)):
        out[f"weights/{name}"] = np.column_stack([np.asarray(itp.x), np.asarray(itp.y)])
    for name, tab in (("brem_elec", d._d_rate_dict_elec_brem), ("brem_positron", d._d_rate_dict_positron_brem),
                      ("muon_brem", d._d_rate_dict_muon_brem), ("annihilation", d._d_rate_dict_positron_ann)):
        keys = list(tab.keys())
        out[f"drate/{name}/E"] = np.array([float(k) for k in keys])
        out[f"drate/{name}/table"] = np.stack([np.asarray(tab[k], dtype=float) for k in keys])
    for P, itp in (("DarkBrem", d._NSigmaDarkBrem), ("DarkAnn", d._NSigmaDarkAnn), ("DarkComp", d._NSigmaDarkComp),
                   ("DarkMuonBrem", d._NSigmaDarkMuonBrem)):
        out[f"nsdark/{P}/x"], out[f"nsdark/{P}/y"] = np.asarray(itp.x), np.asarray(itp.y)
    from petite_b200.tables import mv_tag
    np.savez_compressed(os.path.join(DATA, f"dark_setup_{material}_mV{mv_tag(mV)}.npz"), **out)


def golden_dark(rng):
    from PETITE.dark_shower import DarkShower
    out = {}
    proc_code = {"DarkBrem": 8, "DarkAnn_bound": 9, "DarkComp_bound": 10, "DarkMuonBrem": 11, "TwoBody_BSMDecay": 13}
    for ci, (material, mV) in enumerate(DARK_CASES):
        d = DarkShower(REFDIR, material, 0.010, mV)
        dump_dark_setup(d, material, mV)
        # weight look-ups (GetBSMWeights)
        E = np.geomspace(2e-3, 120.0, 60)
        for pid, pr in ((11, "DarkBrem"), (-11, "DarkBrem"), (-11, "DarkAnn"), (22, "DarkComp"), (13, "DarkMuonBrem"), (11, "DarkAnn")):
            out[f"{ci}/w/{pid}/{pr}"] = np.array([float(d.GetBSMWeights([pid, e], pr)) for e in E])
        out[f"{ci}/w/E"] = E
        pi0 = Particle([5.0, 0, 0, np.sqrt(25 - m_pi0 ** 2)], [0, 0, 0], {"PID": 111, "mass": m_pi0, "stability": "short-lived"})
        out[f"{ci}/w/pi0"] = np.array([float(d.GetBSMWeights(pi0, "TwoBody_BSMDecay"))])
        # whole dark pass on a stream-mode SM shower
        for k, (pid, E0, seed) in enumerate(((11, 5.0, 201), (22, 8.0, 202), (-11, 2.0, 203), (13, 6.0, 204), (111, 6.0, 205))):
            m = {11: m_electron, -11: m_electron, 22: 0.0, 13: m_muon, 111: m_pi0}[pid]
            ids = {"PID": pid, "ID": 1, "mass": m}
            if pid == 111:
                ids["stability"] = "short-lived"
            np.random.seed(seed)
            random.seed(seed)
            sm = d.generate_shower(Particle([E0, 0, 0, np.sqrt(E0 ** 2 - m ** 2)], [0, 0, 0], ids))
            _, vs = d.generate_dark_shower(ExDir=list(sm))
            pre = f"{ci}/sh{k}/"
            out[pre + "case"] = np.array([pid, E0, seed, len(sm)])
            out[pre + "p0"] = np.array([v.get_p0() for v in vs], dtype=float).reshape(-1, 4)
            out[pre + "r0"] = np.array([v.get_r0() for v in vs], dtype=float).reshape(-1, 3)
            out[pre + "weight"] = np.array([v.get_ids()["weight"] for v in vs], dtype=float)
            out[pre + "mass"] = np.array([v.get_ids()["mass"] for v in vs], dtype=float)
            out[pre + "process"] = np.array([proc_code[v.get_ids()["generation_process"]] for v in vs])
            out[pre + "parent_ID_mod"] = np.array([v.get_ids()["parent_ID"] % (1 << 61) for v in vs], dtype=np.int64)
            out[pre + "parent_PID"] = np.array([v.get_ids()["parent_PID"] for v in vs])
            print("dark case", material, mV, pid, E0, "SM", len(sm), "V", len(vs))
        out[f"{ci}/material"] = np.array(material)
        out[f"{ci}/mV"] = np.array(mV)
    out["n_cases"] = np.array(len(DARK_CASES))
    np.savez_compressed(os.path.join(HERE, "dark.npz"), **out)


def golden_dark_kinematics(rng):
    """The three dark-vector kinematics (kinematics.py:43-68, 134-183, 267-299) at points drawn through the shipped dark maps;
    rows = [E, mV, x[4], u1, u2, Pe, cos(theta_e)] -> V four-vector in the parent frame.  The azimuth uniforms the functions
    draw from numpy's global stream are recorded by re-seeding (np.random.uniform(0, 2 pi) = 2 pi * random())."""
    out = {}
    for mV in (0.003, 0.03):
        from petite_b200.tables import mv_tag
        dk = np.load(DATA + f"dark_maps_mV{mv_tag(mV)}.npz")
        for P, pid, mass in (("DarkBrem", 11, m_electron), ("DarkMuonBrem", 13, m_muon), ("DarkAnn", -11, m_electron), ("DarkComp", 22, 0.0)):
            E, ninc, G = dk[f"{P}/E"], dk[f"{P}/ninc"], dk[f"{P}/grid"]
            inp, res = [], []
            for ie in (10, 40, 70, 99):
                grid = split_grid(G[ie], ninc)
                x, _ = map_points(grid, rng.random((24, len(grid))))
                for xi in x:
                    Einc = float(E[ie]) * float(rng.uniform(0.9, 1.0))
                    p = Particle([Einc, 0, 0, np.sqrt(Einc ** 2 - mass ** 2)], [0, 0, 0], {"PID": pid, "mass": mass})
                    seed = int(rng.integers(1 << 30))
                    np.random.seed(seed)
                    u1, u2 = np.random.random(), np.random.random()
                    np.random.seed(seed)
                    Pe, cte = 0.0, 0.0
                    if P in ("DarkBrem", "DarkMuonBrem"):
                        if xi[0] * Einc <= mV:
                            continue
                        v = kin.l_to_lV_fourvecs(p, xi, mV=mV)[1]
                    elif P == "DarkAnn":
                        if 2 * m_electron * (Einc + m_electron) <= mV ** 2:
                            continue
                        v = kin.radiative_return_fourvecs(p, xi, mV=mV)[1]
                    else:
                        Pe, cte = float(1e-3 * rng.random()), float(rng.uniform(-1, 1))
                        ss = m_electron ** 2 + 2 * Einc * (np.sqrt(m_electron ** 2 + Pe ** 2) - cte * Pe)
                        if (ss - mV ** 2 + m_electron ** 2) / (2 * np.sqrt(ss)) < m_electron:
                            continue
                        v = kin.compton_fourvecs_boundelectron(p, xi, mV=mV, Pe=Pe, cte=cte)[1]
                    if not np.all(np.isfinite(np.asarray(v, dtype=float))):
                        continue
                    inp.append([Einc, mV] + list(np.pad(xi, (0, 4 - len(xi)))) + [u1, u2, Pe, cte])
                    res.append(list(np.asarray(v, dtype=float)))
            out[f"{mv_tag(mV)}/{P}/in"] = np.array(inp, dtype=float)
            out[f"{mv_tag(mV)}/{P}/out"] = np.array(res, dtype=float)
            print("dark kinematics", mV, P, len(inp))
    np.savez_compressed(os.path.join(HERE, "dark_kinematics.npz"), **out)


def golden_detector(rng):
    """shower.detector_cut (shower.py:825-864) on a random particle list."""
    from PETITE.shower import detector_cut
    n = 400
    p3 = rng.normal(size=(n, 3)) * np.array([0.05, 0.05, 1.0]) + np.array([0, 0, 2.0])
    E = np.sqrt(np.sum(p3 ** 2, axis=1) + m_electron ** 2)
    r0 = rng.normal(size=(n, 3)) * 0.05
    w = rng.random(n) * 2
    plist = [Particle([E[i], *p3[i]], list(r0[i]), {"PID": 11, "mass": m_electron, "weight": float(w[i])}) for i in range(n)]
    z = [1.0, 10.0, 50.0]
    out = {"p0": np.column_stack([E, p3]), "r0": r0, "w": w, "z": np.array(z)}
    for tag, kw in (("a", dict(detector_radius=0.5)), ("b", dict(detector_radius=2.0, energy_cut=(1.0, 3.0), detector_inner_radius=0.2))):
        out[f"{tag}/total"] = np.array(detector_cut(plist, z, method="TotalWeight", **kw), dtype=float)
        out[f"{tag}/eff"] = np.array(detector_cut(plist, z, method="Efficiency", **kw), dtype=float)
        m = np.array(detector_cut(plist, z, method="SampleW", **kw))
        out[f"{tag}/mask"] = m
        if "energy_cut" in kw:
            sel = (E < kw["energy_cut"][1]) & (E > kw["energy_cut"][0])
            out[f"{tag}/kept"] = np.nonzero(sel)[0]
    np.savez_compressed(os.path.join(HERE, "detector.npz"), **out)


if __name__ == "__main__":
    rng = np.random.default_rng(20261017)
    build_reference_dict_dir()
    if "--detector-only" in sys.argv:
        golden_detector(rng)
        sys.exit(0)
    if "--darkkin-only" in sys.argv:
        golden_dark_kinematics(np.random.default_rng(20261018))
        sys.exit(0)
    if "--dark-only" in sys.argv:
        golden_dark(rng)
        sys.exit(0)
    golden_integrands(rng)
    golden_kinematics(rng)
    golden_mcs(rng)
    golden_particle(rng)
    golden_nsigma()
    golden_showers()
    golden_dark(rng)
    golden_detector(rng)
    golden_dark_kinematics(np.random.default_rng(20261018))
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))
