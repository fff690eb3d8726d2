"""``DarkShower``: PETITE's dark-vector pass in front of the B200 engine.

Same constructor and ``generate_dark_shower`` as the reference (src/PETITE/dark_shower.py:65-129, 806-849); the
per-particle work - weight look-up (GetBSMWeights :595-647), interaction-energy choice, multiple scattering and
energy loss down to it, dark VEGAS accept/reject, dark-vector kinematics (produce_bsm_particle :721-804) - runs as
CUDA kernels (``pb_run_dark``).  The constructor's one-off tables stay on the host: thresholds are computed here,
the quadrature-built weight / dRate tables are read from ``<dict_dir>/dark_setup_<material>_mV<mV>.npz`` (written by
``petite_b200.dark_setup`` or dumped from the reference, tests/golden/make_golden.py) and built on first use otherwise.

Batched entry point: :meth:`DarkShower.generate_dark_showers` takes the :class:`ShowerBatch` still resident in HBM.
"""
import ctypes as C
import os

import numpy as np

from . import _capi as capi
from . import constants as K
from . import tables as tb
from .particle import Particle, meson_twobody_branchingratios
from .shower import Shower, ShowerBatch, LinearTable, stack_struct, new_stack

dark_process_codes = ["DarkBrem", "DarkAnn", "DarkComp", "TwoBody_BSMDecay", "DarkMuonBrem"]
dimensionalities_dark = {"DarkComp": 1, "DarkBrem": 3, "DarkAnn": 1, "DarkMuonBrem": 3}
PID_V = 4900022
_CODE = {"DarkBrem": 8, "DarkAnn": 9, "DarkComp": 10, "DarkMuonBrem": 11, "TwoBody_BSMDecay": 13}
_WEIGHT_ORDER = ("brem_elec", "brem_positron", "annihilation", "muon_brem")


class LogLogTable:
    """``interpolate1d(..., xspace='log', yspace='log', fill_value=-20)`` of dark_shower.py:31-46; nodes are log10."""

    def __init__(self, lx, ly):
        self.x, self.y = np.asarray(lx, dtype=np.float64), np.asarray(ly, dtype=np.float64)
        self._lin = LinearTable(self.x, self.y, fill_value=-20.0)

    def __call__(self, E):
        return 10 ** self._lin(np.log10(E))


class DarkBatch:
    """Dark vectors of one ``pb_run_dark`` call (device-resident stack of PID 4900022 records)."""

    def __init__(self, tensors, n, counters, sm_batch, active):
        self._t, self.n, self.counters, self.sm, self.active = tensors, int(n), counters, sm_batch, list(active)
        self._host = None

    def to_host(self):
        if self._host is None:
            n = self.n
            t = {k: v[:n].cpu().numpy() for k, v in self._t.items() if k != "ids"}
            info = t["meta"][:, 2]
            self._host = dict(p0=t["p0"], r0=t["r0w"][:, :3], weight=t["r0w"][:, 3], parent=t["meta"][:, 1],
                              process=info & 0xFF, shower=t["meta"][:, 3], ntrials=t["aux"][:, 0],
                              generation=(info >> 16) & 0xFFFF)
        return self._host

    def reference_order(self):
        """Dark records ordered as the reference emits them: SM particles in creation order, processes in the
        ``active_processes`` order (dark_shower.py:832-848).  Returns (order, shower_offsets)."""
        h = self.to_host()
        sm_order, _ = self.sm.reference_order()
        rank = np.empty(self.sm.n, dtype=np.int64)
        rank[sm_order] = np.arange(self.sm.n)
        pos = {(_CODE[p]): i for i, p in enumerate(self.active)}
        prank = np.array([pos[int(c)] for c in h["process"]], dtype=np.int64) if self.n else np.zeros(0, dtype=np.int64)
        order = np.lexsort((prank, rank[h["parent"]]))
        counts = np.bincount(h["shower"][order], minlength=self.sm.n_primaries) if self.n else np.zeros(self.sm.n_primaries, int)
        return order, np.concatenate([[0], np.cumsum(counts)])

    def to_particles(self, sm_particles):
        """-> per primary, list of dark-vector :class:`Particle` (ids as dark_shower.py:789-804, 838-842)."""
        h = self.to_host()
        order, offs = self.reference_order()
        sm_order, sm_offs = self.sm.reference_order()
        sm_index = {}
        for i in range(self.sm.n_primaries):
            for k, s in enumerate(sm_order[sm_offs[i]:sm_offs[i + 1]]):
                sm_index[int(s)] = sm_particles[i][k]
        names = {8: "DarkBrem", 9: "DarkAnn_bound", 10: "DarkComp_bound", 11: "DarkMuonBrem", 13: "TwoBody_BSMDecay"}
        out = []
        for i in range(self.sm.n_primaries):
            vs = []
            for j in order[offs[i]:offs[i + 1]]:
                par = sm_index[int(h["parent"][j])].get_ids()
                code = int(h["process"][j])
                ids = {"PID": PID_V, "parent_PID": par["PID"], "parent_ID": par["ID"],
                       "ID": 2 * par["ID"] + (1 if code == 13 else 0), "generation_number": par["generation_number"] + 1,
                       "generation_process": names[code], "weight": float(h["weight"][j])}
                if code == 13:
                    ids["mass"] = self._mV
                vs.append(Particle(np.array(h["p0"][j]), np.array(h["r0"][j]), ids))
            out.append(vs)
        return out


class DarkShower(Shower):
    """A class to reprocess an existing EM shower to generate dark photons (GPU-backed)."""

    def __init__(self, dict_dir, target_material, min_energy, mV_in_GeV, mode="exact", maxF_fudge_global=1,
                 max_n_integrators=int(1e4), kinetic_mixing=1.0, Zeff=29.508, bound_electron=True, g_e=None,
                 active_processes=None, fast_MCS_mode=True, rescale_MCS=1, seed=None, device=None):
        if not bound_electron:
            raise NotImplementedError("bound_electron=False is outside the accelerated scope (the reference's own path "
                                      "raises TypeError for cached materials, dark_shower.py:550)")
        self.active_processes = dark_process_codes if active_processes is None else active_processes
        self.kinetic_mixing = kinetic_mixing
        self.bound_electron = bound_electron
        self.Zeff = Zeff
        self.g_e = kinetic_mixing * np.sqrt(4 * np.pi * K.alpha_em) if g_e is None else g_e
        self._dark_dict_dir = dict_dir
        self._mV_list = tb.list_dark_masses(dict_dir)
        self.set_mV(mV_in_GeV, mode)
        self._dark_ready = False
        super().__init__(dict_dir, target_material, min_energy, maxF_fudge_global=maxF_fudge_global,
                         max_n_integrators=max_n_integrators, fast_MCS_mode=fast_MCS_mode, seed=seed,
                         rescale_MCS=rescale_MCS, device=device)
        self.set_dark_cross_sections()
        self._load_setup()
        self.set_dark_samples()
        self._upload_dark()
        self._dark_stack = None
        self._dark_capacity = 0

    # ------------------------------------------------------------------ reference-compatible set-up
    def get_dark_dict_dir(self):
        return self._dark_dict_dir

    def closest_lesser_value(self, input_list, input_value):
        """dark_shower.py:158-165 incl. the wrap-around below the smallest trained mass (SURVEY Q-2)."""
        arr = np.asarray(input_list)
        index = (np.abs(arr - input_value)).argmin()
        return arr[index] if arr[index] <= input_value else arr[index - 1]

    def set_mV(self, value, mode):
        if mode == "exact":
            self._mV = float(self.closest_lesser_value(self._mV_list, value))
            self._mV_estimator = self._mV
        elif mode == "approx":
            self._mV = value
            self._mV_estimator = float(self.closest_lesser_value(self._mV_list, value))
        else:
            raise Exception("Mode not valid. Chose exact or approx.")

    def get_mV(self):
        return self._mV

    def set_dark_cross_sections(self):
        xs = tb.load_dark_xsec(self._dict_dir, self._mV_estimator, self._target_material)
        for P in ("DarkBrem", "DarkMuonBrem"):
            while xs[P][0][1] == 0.0:          # leading zero rows dropped (dark_shower.py:226-232)
                xs[P] = xs[P][1:]
        self._dark_brem_cross_section, self._dark_annihilation_cross_section = xs["DarkBrem"], xs["DarkAnn"]
        self._dark_compton_cross_section, self._dark_muon_brem_cross_section = xs["DarkComp"], xs["DarkMuonBrem"]
        me = K.m_electron
        self._resonant_annihilation_energy = (self._mV ** 2 - 2 * me ** 2) / (2 * me)
        self._compton_threshold_energy = self._mV ** 2 / (2 * me) + self._mV
        b0, m0 = xs["DarkBrem"][0][0], xs["DarkMuonBrem"][0][0]
        self._minimum_calculable_dark_energy = {
            11: {"DarkBrem": b0}, -11: {"DarkBrem": b0, "DarkAnn": self._resonant_annihilation_energy / 1000.0},
            22: {"DarkComp": self._compton_threshold_energy / 1000.0}, 111: {"TwoBody_BSMDecay": -1},
            # eta, eta': the reference's weight formula covers them (dark_shower.py:633-638) but its threshold table does not
            # (:236-241), so its GetBSMWeights raises KeyError at :604 for these PIDs; the formula is implemented as written
            221: {"TwoBody_BSMDecay": -1}, 331: {"TwoBody_BSMDecay": -1},
            13: {"DarkMuonBrem": m0}, -13: {"DarkMuonBrem": m0}}

    def _setup_path(self):
        return self._dict_dir + f"dark_setup_{self._target_material}_mV{tb.mv_tag(self._mV_estimator)}.npz"

    def _setup_matches(self, path):
        """The cached tables depend on the ACTUAL mV (annihilation / Compton thresholds and bound cross-sections) and on Zeff,
        not only on (material, mV_estimator): the reference recomputes those parts on every construction
        (dark_shower.py:254-275, 311-399).  A cache is used only if its recorded meta equals this object's."""
        try:
            meta = np.load(path)["meta"]
        except Exception:
            return False
        want = (self._mV, self._mV_estimator, self._resonant_annihilation_energy, self._compton_threshold_energy, self.Zeff)
        got = (meta[0], meta[1], meta[2], meta[3], meta[6])
        return all(abs(a - b) <= 1e-12 * max(abs(a), abs(b), 1e-300) for a, b in zip(want, got))

    def _load_setup(self):
        path = self._setup_path()
        cache_dir = os.path.expanduser(os.environ.get("PETITE_B200_CACHE", "~/.cache/petite_b200"))
        exact = self._mV == self._mV_estimator and self.Zeff == 29.508
        name = os.path.basename(path) if exact else os.path.basename(path)[:-4] + f"_m{self._mV:.9g}_Zeff{self.Zeff:.9g}.npz"
        candidates = ([path] if exact else []) + [os.path.join(cache_dir, name)]
        path = next((c for c in candidates if os.path.exists(c) and self._setup_matches(c)), None)
        if path is None:
            # the reference writes its caches into dict_dir and fails on a read-only one (SURVEY Q-3); tables for a
            # non-default (mV, Zeff) and read-only dict_dirs go to a per-user cache directory instead
            from . import dark_setup
            print("Weights not previously calculated, calculating now...")
            try:
                if not exact:
                    raise OSError("non-default mV / Zeff: per-user cache")
                path = dark_setup.build(self, candidates[0])
            except OSError:
                os.makedirs(cache_dir, exist_ok=True)
                path = dark_setup.build(self, candidates[-1])
        z = np.load(path)
        self._weights = {k: LinearTable(z[f"weights/{k}"][:, 0], z[f"weights/{k}"][:, 1]) for k in _WEIGHT_ORDER}
        self._brem_elec_numerical_weight, self._brem_positron_numerical_weight = self._weights["brem_elec"], self._weights["brem_positron"]
        self._annihilation_numerical_weight, self._muon_brem_numerical_weight = self._weights["annihilation"], self._weights["muon_brem"]
        self._drate = {k: (np.ascontiguousarray(z[f"drate/{k}/E"]), np.ascontiguousarray(z[f"drate/{k}/table"])) for k in _WEIGHT_ORDER}
        self._NSigmaDarkComp = LogLogTable(z["nsdark/DarkComp/x"], z["nsdark/DarkComp/y"])
        self._NSigmaDarkBrem = LogLogTable(z["nsdark/DarkBrem/x"], z["nsdark/DarkBrem/y"])
        self._NSigmaDarkAnn = LogLogTable(z["nsdark/DarkAnn/x"], z["nsdark/DarkAnn/y"])
        self._NSigmaDarkMuonBrem = LogLogTable(z["nsdark/DarkMuonBrem/x"], z["nsdark/DarkMuonBrem/y"])

    def set_dark_samples(self):
        procs = [p for p in self.active_processes if p in dimensionalities_dark]
        self._dark_maps = tb.load_dark_maps(self._dict_dir, self._mV_estimator, self._target_material, procs)
        for P, ms in self._dark_maps.items():
            if np.any(np.isnan(ms.max_F)) and os.environ.get("PETITE_B200_ALLOW_MISSING_MAXF"):
                ms.max_F = np.ones(len(ms.E))
            if np.any(np.isnan(ms.max_F)):
                raise Exception(f"no max_F table for dark process {P} / {self._target_material} / mV={self._mV_estimator}")
        self._loaded_dark_samples = {
            P: [[float(ms.E[i]), {"neval": ms.neval, "max_F": {self._target_material: float(ms.max_F[i])},
                                  "adaptive_map": [ms.axis_nodes(i, d) for d in range(ms.dim)]}] for i in range(len(ms.E))]
            for P, ms in self._dark_maps.items()}

    def _config(self):
        c = super()._config()
        c.mV, c.g_e, c.kinetic_mixing, c.Zeff = float(self._mV), float(self.g_e), float(self.kinetic_mixing), float(self.Zeff)
        if hasattr(self, "_resonant_annihilation_energy"):
            c.E_res_ann, c.E_thr_comp = float(self._resonant_annihilation_energy), float(self._compton_threshold_energy)
        c.bound_electron = 1
        return c

    def _upload_dark(self):
        cfg = self._config()
        capi.check(self._engine, capi.lib.pb_set_config(self._engine, C.byref(cfg)))
        for P, ms in self._dark_maps.items():
            self._upload_maps(_CODE[P], ms)
        t = capi.pb_dark_tables()
        keep = []
        for k, name in enumerate(_WEIGHT_ORDER):
            w = self._weights[name]
            E, tab = self._drate[name]
            tab = np.ascontiguousarray(tab, dtype=np.float64)
            keep += [w.x, w.y, E, tab]
            t.w_E[k], t.w_y[k], t.w_n[k] = capi.dptr(w.x), capi.dptr(w.y), len(w.x)
            t.d_E[k], t.d_table[k], t.d_n[k] = capi.dptr(E), capi.dptr(tab), len(E)
        ns = self._NSigmaDarkComp
        t.nsdark_comp_lx, t.nsdark_comp_ly, t.nsdark_comp_n = capi.dptr(ns.x), capi.dptr(ns.y), len(ns.x)
        md = self._minimum_calculable_dark_energy
        for k, v in enumerate((md[11]["DarkBrem"], md[-11]["DarkAnn"], md[22]["DarkComp"], md[13]["DarkMuonBrem"])):
            t.min_E[k] = float(v)
        capi.check(self._engine, capi.lib.pb_upload_dark(self._engine, C.byref(t)))
        self._dark_ready = True

    # ------------------------------------------------------------------ weights (host twin of the device look-up)
    def GetBSMWeights(self, particle, process):
        """dark_shower.py:595-647 for bound_electron=True."""
        if isinstance(particle, (list, np.ndarray)):
            PID, E0 = particle
        else:
            PID, E0 = particle.get_ids()["PID"], particle.get_p0()[0]
        if PID not in [-11, 11, 13, -13, 22, 111, 221, 331]:
            return 0.0
        md = self._minimum_calculable_dark_energy
        if PID not in md or process not in md[PID] or E0 < md[PID][process]:
            return 0.0
        pre = self.g_e ** 2 / (4 * np.pi * K.alpha_em)
        if PID == 22:
            if process != "DarkComp" or E0 < self._minimum_calculable_energy[22]:
                return 0.0
            with np.errstate(all="ignore"):
                return pre * self._NSigmaDarkComp(E0) / (self._NSigmaPP(E0) + self._NSigmaComp(E0))
        if process == "DarkBrem":
            if abs(PID) != 11:
                return 0.0
            return pre * (self._brem_elec_numerical_weight if PID == 11 else self._brem_positron_numerical_weight)(E0)
        if PID == -11 and process == "DarkAnn":
            return pre * self._annihilation_numerical_weight(E0)
        if PID in (111, 221, 331):
            if process != "TwoBody_BSMDecay":
                return 0.0
            r = self._mV / particle.get_ids()["mass"]
            return 0.0 if r >= 1.0 else 2 * self.kinetic_mixing ** 2 * (1.0 - r ** 2) ** 3 * meson_twobody_branchingratios[PID]
        if abs(PID) == 13 and process == "DarkMuonBrem":
            return pre * self._muon_brem_numerical_weight(E0)
        return 0.0

    # ------------------------------------------------------------------ dark pass
    def _ensure_dark_stack(self, capacity):
        if capacity <= self._dark_capacity:
            return
        torch = self._torch
        dev = torch.device("cuda", self._device)
        self._dark_stack = None
        self._dark_stack = new_stack(torch, dev, capacity)
        self._dark_capacity = capacity

    def generate_dark_showers(self, sm_batch):
        """Dark pass over a :class:`ShowerBatch` that is still resident on its stack -> :class:`DarkBatch`.  The batch may
        live on another engine handle's stack (a part of :meth:`Shower.run_arrays_split`): the stack is caller-owned device
        memory, and the dark tables only exist on this object's engine, so the pass always runs here."""
        active = [p for p in self.active_processes if p in _CODE]
        mask = 0
        for p in active:
            mask |= 1 << _CODE[p]
        # dark-stack capacity: the worst case is two dark vectors per SM record (e+: DarkBrem and DarkAnn), showers make 0.8-1.3 per
        # record; start from the larger of the current stack and 1.5 n and fall back to the bound if the engine reports it too small
        # (a DarkBatch of an earlier call keeps ITS stack alive: drop it before the next call if memory is tight)
        worst = 2 * sm_batch.n + 1024
        self._ensure_dark_stack(min(worst, max(self._dark_capacity, int(1.5 * sm_batch.n) + 1024)))
        sm = stack_struct(sm_batch._t)
        cnt = capi.pb_counters()
        stream = self._torch.cuda.current_stream(self._device).cuda_stream
        while True:
            t = self._dark_stack
            dk = stack_struct(t)
            rc = capi.lib.pb_run_dark(self._engine, C.byref(sm), sm_batch.n, mask, C.byref(dk), C.byref(cnt), C.c_void_p(stream))
            if rc == capi.PB_ERR_CAPACITY and self._dark_capacity < worst:
                self._ensure_dark_stack(worst)
                continue
            capi.check(self._engine, rc)
            break
        b = DarkBatch(t, cnt.n_particles, cnt.as_dict(), sm_batch, active)
        b._mV = self._mV
        return b

    def tally_dark(self, dark_batch, out=None):
        """``pb_tally`` over a :class:`DarkBatch` (weighted yield, energy and angle spectra of the dark vectors)."""
        torch = self._torch
        if out is None:
            out = torch.zeros(capi.TALLY_SIZE, dtype=torch.float64, device=torch.device("cuda", self._device))
        st = stack_struct(dark_batch._t)        # the batch's own stack (it may predate a regrow of this object's)
        stream = torch.cuda.current_stream(self._device).cuda_stream
        capi.check(self._engine, capi.lib.pb_tally(self._engine, C.byref(st), 0, dark_batch.n, C.c_void_p(out.data_ptr()),
                                                   C.c_void_p(stream)))
        return out

    def draw_dark_sample(self, Einc, LU_Key=-1, process="DarkBrem", VB=False):
        """One VEGAS accept/reject sample of a dark process at ``Einc`` (dark_shower.py:649-704) -> map variables
        (+ the trial count if ``VB``).  Batched form: ``draw_samples(energies, process)``."""
        if process not in dimensionalities_dark:
            raise Exception("Your process is not in the list")
        if process not in self._dark_maps:
            raise Exception("Process String does not match library")
        x, ntr = self.draw_samples([Einc], process, LU_Key)
        if ntr[0] < 0:
            raise Exception("No Sample Found", process, Einc, LU_Key)
        return np.concatenate([x[0], [ntr[0]]]) if VB else x[0]

    def produce_bsm_particle(self, p0, process, weight=None, VB=False):
        """One dark vector emitted by ``p0`` through ``process`` (dark_shower.py:721-804): interaction-energy choice, multiple
        scattering + energy loss down to it, dark sample, V four-vector in the lab, weight = p0's weight x ``weight`` (default:
        ``GetBSMWeights(p0, process)``, what generate_dark_shower passes).  A one-candidate ``pb_run_dark``.  -> Particle or None."""
        if process not in dimensionalities_dark:
            raise Exception("Your process is not in the list")
        wg = self.GetBSMWeights(p0, process)
        if not wg > 0.0:
            return None
        keep = self.active_processes
        try:
            self.active_processes = [process]
            batch = self.batch_from_particles([p0])
            dark = self.generate_dark_showers(batch)
            vs = dark.to_particles([[p0]])[0]
        finally:
            self.active_processes = keep
        if not vs:
            return None
        v = vs[0]
        if weight is not None:
            v.get_ids()["weight"] = v.get_ids()["weight"] * (float(weight) / wg)
        return v

    def generate_dark_shower(self, ExDir=None, SParams=None):
        """dark_shower.py:806-849: (SM shower, list of dark vectors) for one existing or new SM shower."""
        if ExDir is None and SParams is None:
            print("Need an existing SM shower-file directory or SM incident particle to run dark shower")
            return None
        if ExDir is not None and isinstance(ExDir, str):
            ExDir = list(np.load(ExDir, allow_pickle=True))
        if ExDir is not None and isinstance(ExDir, list):
            batch = self.batch_from_particles(ExDir)
            sm_lists = [ExDir]
        elif isinstance(SParams, Particle):
            batch = self.generate_showers([SParams])
            sm_lists = batch.to_particles([SParams])
        else:
            raise ValueError("Provided SParams must be a `Particle' class object")
        dark = self.generate_dark_showers(batch)
        return sm_lists[0], dark.to_particles(sm_lists)[0]
