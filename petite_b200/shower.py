"""``Shower``: PETITE's SM shower API in front of the B200 engine.

Constructor arguments, attribute names and the per-primary ``generate_shower`` follow the reference
(src/PETITE/shower.py:98-142, 603-708) so existing scripts keep working; the stepping itself runs as CUDA
kernels behind the C ABI of ``include/petite_b200.h`` (there is no CPU path).  What stays on the host is the
one-off set-up the reference also does in Python: reading the tables and tabulating n*sigma(E)
(shower.py:202-295).

New, batched entry point: :meth:`Shower.generate_showers` steps many independent primaries at once and
returns a :class:`ShowerBatch` (structure-of-arrays view of the device stack).
"""
import ctypes as C
import os

import numpy as np

from . import _capi as capi
from . import constants as K
from . import tables as tb
from . import totals
from .particle import Particle, mass_dict

process_code = {"Brem": 0, "Ann": 1, "PairProd": 2, "Comp": 3, "Moller": 4, "Bhabha": 5, "MuonE": 6, "MuonBrem": 7}
dimensionalities = {p: tb.PROC_DIM[p] for p in process_code}
process_PIDS = {"PairProd": [-11, 11], "Brem": [0, 22], "MuonBrem": [0, 22], "Comp": [11, 22], "Ann": [22, 22],
                "Moller": [0, 11], "Bhabha": [0, 11], "MuonE": [0, 11]}
_STEPPING_PIDS = (22, 11, -11, 13, -13)


class LinearTable:
    """1-D table evaluated like ``scipy.interpolate.interp1d(x, y, fill_value=0.0, bounds_error=False)``.

    This is the host-side twin of the device interpolant (``nsigma_eval`` in csrc/engine.cu); the ``x``/``y``
    arrays are exactly what ``pb_upload_nsigma`` ships.
    """

    def __init__(self, x, y, fill_value=0.0):
        self.x = np.ascontiguousarray(x, dtype=np.float64)
        self.y = np.ascontiguousarray(y, dtype=np.float64)
        self.fill_value = fill_value

    def __call__(self, xn):
        xn_arr = np.asarray(xn, dtype=np.float64)
        x, y = self.x, self.y
        hi = np.clip(np.searchsorted(x, xn_arr), 1, len(x) - 1)
        lo = hi - 1
        slope = (y[hi] - y[lo]) / (x[hi] - x[lo])
        out = slope * (xn_arr - x[lo]) + y[lo]
        out = np.where((xn_arr < x[0]) | (xn_arr > x[-1]), self.fill_value, out)
        return out if out.ndim else np.float64(out)


def stack_struct(t):
    """``pb_stack`` over a dict of stack tensors (capacity = their length)."""
    return capi.pb_stack(t["p0"].data_ptr(), t["r0w"].data_ptr(), t["pf"].data_ptr(), t["rf"].data_ptr(),
                         t["ids"].data_ptr(), t["aux"].data_ptr(), int(t["p0"].shape[0]))


def new_stack(torch, dev, capacity):
    """Stack tensors for ``capacity`` records.  ``meta`` (pid, parent, info, shower) and ``key`` are views into the packed
    32-byte ``ids`` records (include/petite_b200.h)."""
    f64 = lambda: torch.empty((capacity, 4), dtype=torch.float64, device=dev)
    ids = torch.empty((capacity, 8), dtype=torch.int32, device=dev)
    return {"p0": f64(), "r0w": f64(), "pf": f64(), "rf": f64(), "ids": ids, "meta": ids[:, 0:4], "key": ids[:, 4:6],
            "aux": torch.zeros((capacity, 2), dtype=torch.int32, device=dev)}


class ShowerBatch:
    """Result of a batched run: the filled part of the particle stack, still on the GPU (torch tensors).

    Columns: ``p0`` (n,4), ``r0`` (n,3), ``weight`` (n,), ``pf`` (n,4), ``rf`` (n,3), ``pid``, ``parent`` (slot of the
    parent record, -1 for primaries), ``generation``, ``child_bit``, ``process`` (pb_process code), ``flags``,
    ``shower`` (index of the primary within the call), ``ntrials``, ``nsub``.
    """

    def __init__(self, owner, tensors, n, counters, n_primaries, first_shower_id):
        self._owner = owner
        self._t = tensors
        self.n = int(n)
        self.counters = counters
        self.n_primaries = int(n_primaries)
        self.first_shower_id = int(first_shower_id)
        self._host = None
        # a batch is a VIEW of its owner's reusable stack: the next run on the same Shower overwrites it (copy what is needed with
        # to_host() / to_particles() first, or step sub-batches on their own handles: run_arrays_split)
        self._serial = getattr(owner, "_run_serial", 0)

    def _check_live(self, device=False):
        o = self._owner
        if (device or self._host is None) and getattr(o, "_stack_tensors", None) is self._t and getattr(o, "_run_serial", 0) != self._serial:
            import warnings
            warnings.warn("this ShowerBatch was overwritten by a later run on the same Shower: its device records are those of the "
                          "newer batch (call to_host() before the next run, or use run_arrays_split)", RuntimeWarning, stacklevel=3)

    def device(self, name):
        """Raw device column (torch tensor view over the first ``n`` records)."""
        t = self._t[name]
        return t[: self.n]

    def to_host(self):
        self._check_live()
        if self._host is None:
            n = self.n
            t = {k: v[:n].cpu().numpy() for k, v in self._t.items() if k != "ids"}
            meta = t["meta"]
            info = meta[:, 2]
            self._host = dict(
                p0=t["p0"], r0=t["r0w"][:, :3], weight=t["r0w"][:, 3], pf=t["pf"], rf=t["rf"][:, :3], mass=t["rf"][:, 3],
                pid=meta[:, 0], parent=meta[:, 1], generation=(info >> 16) & 0xFFFF, child_bit=(info >> 15) & 1,
                flags=(info >> 8) & 0x7F, process=info & 0xFF, shower=meta[:, 3],
                ntrials=t["aux"][:, 0], nsub=t["aux"][:, 1])
        return self._host

    def reference_order(self):
        """Permutation of record slots into the reference's creation order, shower by shower.

        ``generate_shower`` appends daughters while iterating over the growing list (shower.py:634-706), so its
        output is ordered by wave, then by the parent's position, then first/second daughter.  Records are stored
        wave by wave on the GPU but unordered inside a wave; this rebuilds the order from (parent, child_bit).
        Returns ``(order, shower_offsets)``: ``order[shower_offsets[i]:shower_offsets[i+1]]`` are shower i's slots.
        """
        h = self.to_host()
        n = self.n
        gen = h["generation"].astype(np.int64)
        gen0 = gen[: self.n_primaries]
        depth = gen - gen0[h["shower"]]
        rank = np.zeros(n, dtype=np.int64)
        rank[: self.n_primaries] = np.arange(self.n_primaries)
        wave_order = [np.arange(self.n_primaries)]
        # generation by generation (slots are appended wave by wave, but a track whose sub-step loop was carried over several waves
        # - engine.cu, k_loop - emits its daughters later than its generation's wave: group by depth explicitly)
        by_depth = np.argsort(depth, kind="stable")
        bounds = np.searchsorted(depth[by_depth], np.arange(1, int(depth.max()) + 2 if n else 1))
        b = self.n_primaries
        for e in bounds:
            if e <= b:
                continue
            sl = by_depth[b:e]
            key = rank[h["parent"][sl]] * 2 + h["child_bit"][sl]
            o = np.argsort(key, kind="stable")
            rank[sl[o]] = np.arange(len(sl))
            wave_order.append(sl[o])
            b = e
        allslots = np.concatenate(wave_order)
        o = np.argsort(h["shower"][allslots], kind="stable")
        order = allslots[o]
        counts = np.bincount(h["shower"][order], minlength=self.n_primaries)
        return order, np.concatenate([[0], np.cumsum(counts)])

    def to_particles(self, primaries=None):
        """-> list (one per primary) of lists of :class:`Particle` in the reference's order and conventions."""
        h = self.to_host()
        order, offs = self.reference_order()
        ref_ids = {}
        out = []
        for i in range(self.n_primaries):
            plist = []
            for s in order[offs[i]:offs[i + 1]]:
                s = int(s)
                pid = int(h["pid"][s])
                par = int(h["parent"][s])
                proc = K.PROCESS_NAMES[int(h["process"][s])]
                if par < 0:
                    src = primaries[i].get_ids() if primaries is not None else {}
                    ids = dict(src)
                    ids.setdefault("PID", pid)
                    ids.setdefault("ID", 1)
                    ids["mass"] = float(h["mass"][s])
                else:
                    pids = ref_ids[par]
                    ids = {"PID": pid, "ID": 2 * pids["ID"] + int(h["child_bit"][s]),
                           "generation_number": pids["generation_number"] + 1, "generation_process": proc,
                           "weight": float(h["weight"][s]), "mass": mass_dict[pid]}
                    if proc == "SMDecay":
                        # decay daughters keep the default parent ids (particle.py:407-408; SURVEY Q-10)
                        ids["production_time"] = pids.get("decay_time", 0.0)
                    else:
                        ids["parent_PID"] = pids["PID"]
                        ids["parent_ID"] = pids["ID"]
                p = Particle(np.array(h["p0"][s]), np.array(h["r0"][s]), ids)
                p.set_pf(np.array(h["pf"][s]))
                p.set_rf(np.array(h["rf"][s]))
                p.set_ended(True)
                ref_ids[s] = p.get_ids()
                plist.append(p)
            out.append(plist)
        return out


class Shower:
    """Representation of a shower (GPU-backed).  Same constructor as the reference (shower.py:100-110)."""

    def __init__(self, dict_dir, target_material, min_energy, maxF_fudge_global=1, max_n_integrators=int(1e4),
                 fast_MCS_mode=True, seed=None, rescale_MCS=1, device=None):
        if not fast_MCS_mode:
            raise NotImplementedError("Bethe-Moliere multiple scattering is outside the accelerated path "
                                      "(and raises TypeError in the reference, moliere.py:244)")
        import torch  # device memory + streams only
        if not torch.cuda.is_available():
            raise RuntimeError("petite_b200 needs a CUDA device: the shower path has no CPU fallback")
        self._torch = torch
        self._device = torch.cuda.current_device() if device is None else int(device)
        # seed=None: fresh entropy, as the reference leaves NumPy's global generator unseeded (shower.py:140-142); the
        # value actually used is exposed as ``self.seed`` so that a run can be reproduced
        self._seed = int.from_bytes(os.urandom(8), "little") if seed is None else int(seed) & 0xFFFFFFFFFFFFFFFF
        self.seed = self._seed
        self._next_shower_id = 0
        self.set_dict_dir(dict_dir)
        self.set_target_material(target_material)
        self.min_energy = min_energy
        self.set_material_properties()
        self.set_n_targets()
        self.set_cross_sections()
        self.set_samples()
        self.set_NSigmas()
        self._MCS_rescale_factor = rescale_MCS
        self._maxF_fudge_global = maxF_fudge_global
        self._max_n_integrators = max_n_integrators
        self._engine = capi.pb_engine()
        self._stack_tensors = None
        self._stack_capacity = 0
        self._create_engine()

    # ------------------------------------------------------------------ reference-compatible set-up
    def set_dict_dir(self, value):
        self._dict_dir = value

    def get_dict_dir(self):
        return self._dict_dir

    def set_target_material(self, value):
        self._target_material = value

    def get_target_material(self):
        return self._target_material

    def set_material_properties(self):
        info = K.target_information[self.get_target_material()]
        self._ZTarget, self._ATarget, self._rhoTarget, self._dEdx = info["Z_T"], info["A_T"], info["rho"], info["dEdx"]

    def get_material_properties(self):
        return self._ZTarget, self._ATarget, self._rhoTarget, self._dEdx

    def set_n_targets(self):
        ZT, AT, rhoT, _ = self.get_material_properties()
        self._nTarget = rhoT / K.m_proton_grams / AT
        self._nElecs = self._nTarget * ZT

    def get_n_targets(self):
        return self._nTarget, self._nElecs

    def set_cross_sections(self):
        xs = tb.load_sm_xsec(self._dict_dir, self._target_material)
        mue = xs["MuonE"]
        while mue[0][1] == 0.0:          # leading zero rows are dropped (shower.py:238-239)
            mue = mue[1:]
        xs["MuonE"] = mue
        self._xsec = xs
        self._brem_cross_section, self._pair_production_cross_section = xs["Brem"], xs["PairProd"]
        self._annihilation_cross_section, self._compton_cross_section = xs["Ann"], xs["Comp"]
        self._moller_cross_section, self._bhabha_cross_section = xs["Moller"], xs["Bhabha"]
        self._muonbrem_cross_section, self._muone_cross_section = xs["MuonBrem"], xs["MuonE"]
        lo = {p: xs[p][0][0] for p in xs}
        mu_lo = max(min(lo["MuonBrem"], lo["MuonE"]), 0.120)
        self._minimum_calculable_energy = {11: min(lo["Brem"], lo["Moller"]),
                                           -11: min(lo["Brem"], lo["Bhabha"], lo["Ann"]),
                                           13: mu_lo, -13: mu_lo, 22: min(lo["PairProd"], lo["Comp"])}

    def get_brem_cross_section(self):
        return self._brem_cross_section

    def get_pairprod_cross_section(self):
        return self._pair_production_cross_section

    def get_annihilation_cross_section(self):
        return self._annihilation_cross_section

    def get_compton_cross_section(self):
        return self._compton_cross_section

    def get_moller_cross_section(self):
        return self._moller_cross_section

    def get_bhabha_cross_section(self):
        return self._bhabha_cross_section

    def get_muonbrem_cross_section(self):
        return self._muonbrem_cross_section

    def get_muone_cross_section(self):
        return self._muone_cross_section

    def set_samples(self):
        self._maps = tb.load_sm_maps(self._dict_dir, self._target_material)
        import os
        for P, ms in self._maps.items():
            if np.any(np.isnan(ms.max_F)):
                if os.environ.get("PETITE_B200_ALLOW_MISSING_MAXF"):      # bootstrap: tables about to be built with find_max
                    ms.max_F = np.ones(len(ms.E))
                    continue
                raise Exception(f"no max_F table for process {P} / material {self._target_material} in {self._dict_dir}")
        # reference-shaped view: _loaded_samples[process][i] = [E_inc, {...}] (shower.py:210-215)
        self._loaded_samples = {
            P: [[float(ms.E[i]), {"neval": ms.neval, "max_F": {self._target_material: float(ms.max_F[i])},
                                  "adaptive_map": [ms.axis_nodes(i, d) for d in range(ms.dim)],
                                  "Eg_min": ms.Eg_min, "Ee_min": ms.Ee_min}] for i in range(len(ms.E))]
            for P, ms in self._maps.items()}
        self._Egamma_min = self._maps["Brem"].Eg_min
        self._Ee_min = self._maps["Brem"].Ee_min

    def set_NSigmas(self):
        """n*sigma(E) in 1/cm for the 8 processes (shower.py:273-295): tables for Brem/PairProd/Ann/Comp/MuonBrem,
        closed forms on a geometric grid for Moller/Bhabha/MuonE."""
        X = self._xsec
        nZ, ne = self.get_n_targets()
        G = K.GeVsqcm2
        t = {}
        for P, n in (("Brem", nZ), ("PairProd", nZ), ("Ann", ne), ("Comp", ne), ("MuonBrem", nZ)):
            t[P] = LinearTable(X[P][:, 0], n * G * X[P][:, 1])
        BS = X["Brem"]
        ee_grid = np.geomspace(3.0 * K.m_electron + self._Ee_min, BS[-1][0], len(BS))
        t["Moller"] = LinearTable(ee_grid, ne * G * totals.sigma_moller(ee_grid, self._Ee_min))
        t["Bhabha"] = LinearTable(ee_grid, ne * G * totals.sigma_bhabha(ee_grid, self._Ee_min))
        self._muon_e_minimum = max(float(totals.muone_threshold(self._Ee_min)), X["MuonE"][0][0])
        mu_grid = np.geomspace(self._muon_e_minimum, ee_grid[-1], len(BS))
        t["MuonE"] = LinearTable(mu_grid, ne * G * totals.sigma_muone(mu_grid, self._Ee_min))
        self._nsigma_tables = t
        self._NSigmaBrem, self._NSigmaPP, self._NSigmaAnn, self._NSigmaComp = t["Brem"], t["PairProd"], t["Ann"], t["Comp"]
        self._NSigmaMoller, self._NSigmaBhabha = t["Moller"], t["Bhabha"]
        self._NSigmaMuonE, self._NSigmaMuonBrem = t["MuonE"], t["MuonBrem"]

    def _NSigmaElectron(self, E):
        return self._NSigmaBrem(E) + self._NSigmaMoller(E)

    def _NSigmaPhoton(self, E):
        return self._NSigmaPP(E) + self._NSigmaComp(E)

    def _NSigmaPositron(self, E):
        return self._NSigmaBrem(E) + self._NSigmaBhabha(E) + self._NSigmaAnn(E)

    def _NSigmaMuon(self, E):
        return self._NSigmaMuonBrem(E) + self._NSigmaMuonE(E)

    def get_mfp(self, particle):
        """Mean free path in metres (shower.py:370-389)."""
        if not isinstance(particle, Particle) and isinstance(particle, (list, np.ndarray)):
            PID, Energy = particle
        else:
            PID, Energy = particle.get_ids()["PID"], particle.get_pf()[0]
        if PID == 22:
            ns = self._NSigmaPhoton(Energy)
        elif PID == 11:
            ns = self._NSigmaElectron(Energy)
        elif PID == -11:
            ns = self._NSigmaPositron(Energy)
        elif abs(PID) == 13:
            ns = self._NSigmaMuon(Energy)
        if ns <= 0.0:
            return 1.0e12
        return K.cmtom / ns

    def BF_positron_brem(self, Energy):
        b0, b1 = self._NSigmaBrem(Energy), self._NSigmaAnn(Energy)
        return b0 / (b0 + b1)

    def BF_photon_pairprod(self, Energy):
        b0, b1 = self._NSigmaPP(Energy), self._NSigmaComp(Energy)
        return b0 / (b0 + b1)

    # ------------------------------------------------------------------ engine plumbing
    def _config(self):
        c = capi.pb_config()
        c.Z_T, c.A_T, c.rho = float(self._ZTarget), float(self._ATarget), float(self._rhoTarget)
        c.dEdx_GeV_per_m = self._dEdx * 0.1
        c.mT_sampler = float(self._ATarget)          # event_info['mT'] = A_T at sampling time (shower.py:435)
        c.min_energy = float(self.min_energy)
        c.Eg_min, c.Ee_min = float(self._Egamma_min), float(self._Ee_min)
        c.maxF_fudge = float(self._maxF_fudge_global)
        c.rescale_MCS = float(self._MCS_rescale_factor)
        for i, pid in enumerate((11, -11, 22, 13, -13)):
            c.min_calc[i] = float(self._minimum_calculable_energy[pid])
        c.max_sweeps = int(self._max_n_integrators)
        return c

    def _create_engine(self):
        cfg = self._config()
        rc = capi.lib.pb_create(C.byref(self._engine), self._device, C.byref(cfg))
        if rc != capi.PB_OK:
            raise capi.EngineError(rc, "pb_create failed (is a CUDA device visible?)")
        for P, code in process_code.items():
            t = self._nsigma_tables[P]
            capi.check(self._engine, capi.lib.pb_upload_nsigma(self._engine, code, capi.dptr(t.x), capi.dptr(t.y), len(t.x)))
            self._upload_maps(code, self._maps[P])

    def _upload_maps(self, code, ms):
        grid = np.ascontiguousarray(ms.grid, dtype=np.float64)
        ninc = np.ascontiguousarray(ms.ninc, dtype=np.int32)
        E = np.ascontiguousarray(ms.E, dtype=np.float64)
        mf = np.ascontiguousarray(ms.max_F, dtype=np.float64)
        capi.check(self._engine, capi.lib.pb_upload_maps(self._engine, code, capi.dptr(grid), len(E), ms.dim, capi.iptr(ninc),
                                                         capi.dptr(E), capi.dptr(mf), int(ms.neval)))

    def __del__(self):
        try:
            if getattr(self, "_engine", None):
                capi.lib.pb_destroy(self._engine)
                self._engine = None
        except Exception:
            pass

    def _ensure_stack(self, capacity):
        if capacity <= self._stack_capacity:
            return
        torch = self._torch
        dev = torch.device("cuda", self._device)
        self._stack_tensors = None   # release before allocating the bigger one
        self._stack_tensors = new_stack(torch, dev, capacity)
        self._stack_capacity = capacity

    def estimate_records(self, energies, pids):
        """Stack records needed for a batch.  Small batches use a generous closed-form bound; large ones are sized from
        a pilot run of 512 of their own primaries (records per GeV and widest wave), so that HBM is not over-allocated.
        The engine fails loudly (PB_ERR_CAPACITY) if the estimate is exceeded and ``run_arrays`` retries once, doubled."""
        E = np.asarray(energies, dtype=np.float64)
        bound = int(np.sum(64 + 4.0 * E / max(self.min_energy, 1e-4)) * 1.5) + 1024
        if len(E) <= 2048 or bound < (1 << 22):
            return bound
        key = (len(E), float(E.sum()))
        if getattr(self, "_pilot_key", None) != key:
            self._pilot_key, self._pilot = key, None
        return bound if self._pilot is None else min(bound, self._pilot)

    def _pilot_capacity(self, p, r, w, m, pid, flags, GlobalMS):
        n = len(pid)
        idx = np.linspace(0, n - 1, 512).astype(np.int64)
        b = self.run_arrays(p[idx], r[idx], w[idx], m[idx], pid[idx], flags[idx], GlobalMS=GlobalMS,
                            capacity=int(np.sum(64 + 4.0 * p[idx, 0] / max(self.min_energy, 1e-4)) * 1.5) + 1024,
                            first_shower_id=(1 << 62))
        scale = float(np.sum(p[:, 0]) / max(np.sum(p[idx, 0]), 1e-300))
        return int(b.n * scale * 1.10 + 2.5 * b.counters["max_wave"] * scale) + (1 << 16)

    @staticmethod
    def _pack_primaries(plist):
        n = len(plist)
        p = np.empty((n, 4)); r = np.empty((n, 3)); w = np.empty(n); m = np.empty(n)
        pid = np.empty(n, dtype=np.int32); fl = np.zeros(n, dtype=np.int32)
        for i, q in enumerate(plist):
            ids = q.get_ids()
            p[i] = np.asarray(q.get_p0(), dtype=np.float64)
            r[i] = np.asarray(q.get_r0(), dtype=np.float64)
            w[i] = ids["weight"]; m[i] = ids["mass"]; pid[i] = ids["PID"]
            st = ids["stability"]
            if st == "short-lived":
                if ids["PID"] != 111:
                    # eta / eta' pick a three-body channel at random and raise, omega's pi0 daughter is never ended (particle.py:391-409)
                    raise ValueError("only short-lived pi0 -> gamma gamma decays are handled on the GPU path")
                fl[i] = capi.PB_FLAG_SHORT_LIVED
            elif st == "long-lived":
                if abs(ids["PID"]) not in (211, 321):
                    raise ValueError("Decay options for particle not specified.")      # particle.py:392-393 / int_length_dict
                fl[i] = capi.PB_FLAG_LONG_LIVED                                        # pi+-, K+- -> mu nu in flight (particle.py:410-422)
            elif ids["PID"] not in _STEPPING_PIDS and abs(ids["PID"]) != 14:
                # the reference never ends such a particle and loops forever (SURVEY Q-20)
                raise ValueError(f"stable PID {ids['PID']} cannot be showered")
        return p, r, w, m, pid, fl

    def _stack_struct(self):
        return stack_struct(self._stack_tensors)

    def run_arrays(self, p, r, w, m, pid, flags, GlobalMS=True, capacity=None, first_shower_id=None):
        """Lowest-level entry: SoA primaries -> :class:`ShowerBatch` (one ``pb_run_showers`` call).

        The six arrays are either all host NumPy arrays (copied to the GPU inside the call) or all torch CUDA
        tensors already resident in HBM (float64 / int32, contiguous)."""
        n = len(pid)
        if n == 0:
            self._ensure_stack(1024)
            return ShowerBatch(self, self._stack_tensors, 0, capi.pb_counters().as_dict(), 0, first_shower_id or 0)
        on_device = not isinstance(p, np.ndarray)
        if on_device:
            arrs = [p, r, w, m, pid, flags]
            assert all(a.is_cuda and a.is_contiguous() for a in arrs)
            if capacity is None:
                raise ValueError("capacity must be given for device-resident primaries")
            cast = lambda a, ty: C.cast(C.c_void_p(a.data_ptr()), ty)
            prim = capi.pb_primaries(cast(p, capi.c_double_p), cast(r, capi.c_double_p), cast(w, capi.c_double_p),
                                     cast(m, capi.c_double_p), cast(pid, capi.c_int32_p), cast(flags, capi.c_int32_p), n, 1)
        else:
            p = np.ascontiguousarray(p, dtype=np.float64); r = np.ascontiguousarray(r, dtype=np.float64)
            w = np.ascontiguousarray(w, dtype=np.float64); m = np.ascontiguousarray(m, dtype=np.float64)
            pid = np.ascontiguousarray(pid, dtype=np.int32); flags = np.ascontiguousarray(flags, dtype=np.int32)
            auto = capacity is None
            if auto:
                capacity = self.estimate_records(p[:, 0], pid)
                if n > 2048 and capacity >= (1 << 22) and getattr(self, "_pilot", None) is None:
                    self._pilot = self._pilot_capacity(p, r, w, m, pid, flags, GlobalMS)
                    capacity = min(capacity, self._pilot)
            prim = capi.pb_primaries(capi.dptr(p), capi.dptr(r), capi.dptr(w), capi.dptr(m), capi.iptr(pid), capi.iptr(flags), n, 0)
        self._ensure_stack(int(capacity))
        if first_shower_id is None:
            first_shower_id = self._next_shower_id
            self._next_shower_id += n
        t = self._stack_tensors
        st = self._stack_struct()
        cnt = capi.pb_counters()
        stream = self._torch.cuda.current_stream(self._device).cuda_stream
        self._run_serial = getattr(self, "_run_serial", 0) + 1           # earlier batches of this stack are stale from here on
        rc = capi.lib.pb_run_showers(self._engine, C.byref(prim), self._seed, int(first_shower_id), 1 if GlobalMS else 0,
                                     C.byref(st), C.byref(cnt), C.c_void_p(stream))
        if rc == capi.PB_ERR_CAPACITY and not on_device and auto:
            # pilot estimate too small (heavy-tailed batch): one retry with twice the room; showers depend only on
            # (seed, shower id), so the rerun reproduces the same particles
            self._pilot = None
            return self.run_arrays(p, r, w, m, pid, flags, GlobalMS=GlobalMS, capacity=2 * int(capacity),
                                   first_shower_id=first_shower_id)
        capi.check(self._engine, rc)
        return ShowerBatch(self, t, cnt.n_particles, cnt.as_dict(), n, first_shower_id)

    # ------------------------------------------------------------------ concurrent sub-batches
    def _clone_engine(self):
        """A second engine handle on the same GPU with the same tables (own scratch, own stack).  The host-side tables
        are shared by reference; only the device copies (13.5 MB) are duplicated."""
        import copy
        peer = copy.copy(self)
        peer._peers, peer._pool, peer._peer_streams = [], None, []
        peer._engine = capi.pb_engine()
        peer._stack_tensors, peer._stack_capacity = None, 0
        peer._create_engine()
        return peer

    def _ensure_peers(self, parts):
        from concurrent.futures import ThreadPoolExecutor
        torch = self._torch
        peers = self.__dict__.setdefault("_peers", [])
        streams = self.__dict__.setdefault("_peer_streams", [])
        with torch.cuda.device(self._device):
            while len(peers) < parts - 1:
                peers.append(self._clone_engine())
            while len(streams) < parts:
                streams.append(torch.cuda.Stream(device=self._device))
        if self.__dict__.get("_pool") is None or self._pool._max_workers < parts:
            self._pool = ThreadPoolExecutor(max_workers=parts)
        return [self] + peers[: parts - 1], streams[:parts]

    def run_arrays_split(self, p, r, w, m, pid, flags, parts=2, GlobalMS=True, capacity=None, first_shower_id=None, tally=None):
        """``run_arrays`` over ``parts`` contiguous sub-batches stepped CONCURRENTLY: one engine handle, one stack, one
        CUDA stream and one host thread per part (handles are independent, include/petite_b200.h).  Every wave kernel
        is a persistent grid whose last CTAs finish late (the longest track / tile of the wave) and the shrinking tail
        of a batch is latency-bound; a second batch in flight fills both.  Showers are keyed by (seed, shower id) and
        part k starts at ``first_shower_id + offset_k``, so the particles are those of the single-batch call.

        ``tally`` (torch float64[TALLY_SIZE] on this GPU, zeroed by the caller): every part adds its ``pb_tally`` to it on
        its own stream as soon as it is done, i.e. while the other part is still in its narrow last waves.

        Stream semantics are those of a synchronous call on the caller's current stream: the parts start after the
        work already queued there and the current stream waits for all of them.  -> list of :class:`ShowerBatch`."""
        torch = self._torch
        n = len(pid)
        parts = max(1, min(int(parts), n))
        if first_shower_id is None:
            first_shower_id = self._next_shower_id
            self._next_shower_id += n
        if parts == 1:
            b = self.run_arrays(p, r, w, m, pid, flags, GlobalMS=GlobalMS, capacity=capacity, first_shower_id=first_shower_id)
            if tally is not None:
                self.tally(b, tally)
            return [b]
        engines, streams = self._ensure_peers(parts)
        from .distributed import shard             # the same contiguous partition as across ranks (tests/test_distributed_cpu.py)
        bounds = [shard(n, k, parts)[0] for k in range(parts)] + [n]
        cap = None if capacity is None else int(capacity) // parts + (1 << 16)
        cur = torch.cuda.current_stream(self._device)
        start = torch.cuda.Event()
        start.record(cur)

        def work(k):
            sl = slice(bounds[k], bounds[k + 1])
            with torch.cuda.device(self._device), torch.cuda.stream(streams[k]):
                streams[k].wait_event(start)
                b = engines[k].run_arrays(p[sl], r[sl], w[sl], m[sl], pid[sl], flags[sl], GlobalMS=GlobalMS, capacity=cap,
                                          first_shower_id=first_shower_id + bounds[k])
                if tally is not None:
                    engines[k].tally(b, tally)
                done = torch.cuda.Event()
                done.record(streams[k])
            return b, done

        out = [f.result() for f in [self._pool.submit(work, k) for k in range(parts)]]
        for _, done in out:
            cur.wait_event(done)
        return [b for b, _ in out]

    def tally_batches(self, batches, out=None):
        """``tally`` over the batches of a split run, accumulated into one buffer."""
        for b in batches:
            out = b._owner.tally(b, out)
        return out

    def run_tallies(self, p, r, w, m, pid, flags, batch=32768, GlobalMS=True, first_shower_id=0, dark=None):
        """Memory-bounded run: step the primaries ``batch`` at a time, keep only the tallies (``pb_tally``), discard the
        particle history.  ``dark`` (a DarkShower sharing this object) adds the dark-vector tallies of each batch.
        Returns (sm_tally, dark_tally or None, summed counters)."""
        torch = self._torch
        dev = torch.device("cuda", self._device)
        sm_t = torch.zeros(capi.TALLY_SIZE, dtype=torch.float64, device=dev)
        dk_t = torch.zeros(capi.TALLY_SIZE, dtype=torch.float64, device=dev) if dark is not None else None
        tot = {}
        n = len(pid)
        for a in range(0, n, batch):
            sl = slice(a, min(a + batch, n))
            b = self.run_arrays(p[sl], r[sl], w[sl], m[sl], pid[sl], flags[sl], GlobalMS=GlobalMS, first_shower_id=first_shower_id + a)
            self.tally(b, sm_t)
            for k, v in b.counters.items():
                tot[k] = max(tot.get(k, 0), v) if k == "max_wave" else tot.get(k, 0) + v
            if dark is not None:
                d = dark.generate_dark_showers(b)
                dark.tally_dark(d, dk_t)
                tot["n_dark"] = tot.get("n_dark", 0) + d.n
        return sm_t, dk_t, tot

    def tally(self, batch, out=None):
        """Histogram / yield tallies of a batch (``pb_tally``), accumulated into ``out`` (torch float64[1024], CUDA)."""
        torch = self._torch
        if out is None:
            out = torch.zeros(capi.TALLY_SIZE, dtype=torch.float64, device=torch.device("cuda", self._device))
        batch._check_live(device=True)
        st = stack_struct(batch._t)             # the batch's own stack (it may predate a regrow of this object's)
        stream = torch.cuda.current_stream(self._device).cuda_stream
        capi.check(self._engine, capi.lib.pb_tally(self._engine, C.byref(st), 0, batch.n, C.c_void_p(out.data_ptr()),
                                                   C.c_void_p(stream)))
        return out

    def set_profiling(self, level=2):
        """0 off; 1 = CUDA-event timing of the two dominant kernels only; 2 = every kernel (adds ~6 % to a step)."""
        capi.check(self._engine, capi.lib.pb_set_profiling(self._engine, int(level)))

    def get_profile(self):
        """Per-kernel device milliseconds / launches and per-process trial counts of the last run."""
        pr = capi.pb_profile()
        capi.check(self._engine, capi.lib.pb_get_profile(self._engine, C.byref(pr)))
        return {"ms": {k: pr.ms[i] for i, k in enumerate(capi.KERNEL_NAMES)},
                "launches": {k: int(pr.launches[i]) for i, k in enumerate(capi.KERNEL_NAMES)},
                "trials": {K.PROCESS_NAMES[i]: int(pr.trials[i]) for i in range(12) if pr.trials[i]},
                "samples": {K.PROCESS_NAMES[i]: int(pr.samples[i]) for i in range(12) if pr.samples[i]}}

    def measure_fp64_peak(self):
        v = (C.c_double * 1)(0.0)
        capi.check(self._engine, capi.lib.pb_measure_fp64_peak(self._engine, v))
        return float(v[0])

    # ------------------------------------------------------------------ single-process sampling (tutorial API)
    _ALL_CODES = dict(process_code, DarkBrem=8, DarkAnn=9, DarkComp=10, DarkMuonBrem=11)

    def draw_samples(self, Einc, process, LU_Key=-1, first_id=None):
        """Batched ``draw_sample``: energies (n,) -> (x (n, dim), trials (n,)).  Sample i uses the Philox key
        (seed, first_id + i); ``trials`` is the reference's ``VB`` counter."""
        E = np.ascontiguousarray(np.atleast_1d(Einc), dtype=np.float64)
        n = len(E)
        if first_id is None:
            first_id = self._next_shower_id
            self._next_shower_id += n
        x = np.zeros((n, 4))
        ntr = np.zeros(n, dtype=np.int32)
        stream = self._torch.cuda.current_stream(self._device).cuda_stream
        rc = capi.lib.pb_draw_samples(self._engine, self._ALL_CODES[process], capi.dptr(E), n, int(LU_Key), self._seed,
                                      int(first_id), capi.dptr(x), capi.iptr(ntr), C.c_void_p(stream))
        capi.check(self._engine, rc)
        return x[:, :tb.PROC_DIM[process]], ntr

    def draw_sample(self, Einc, LU_Key=-1, process='PairProd', VB=False):
        """One VEGAS accept/reject sample for ``process`` at ``Einc`` (shower.py:401-465)."""
        if process not in process_code:
            raise Exception("Your process is not in the list")
        x, ntr = self.draw_samples([Einc], process, LU_Key)
        if ntr[0] < 0:
            raise Exception("No Sample Found", process, Einc, LU_Key)
        return np.concatenate([x[0], [ntr[0]]]) if VB else x[0]

    def sample_scattering(self, p0, process, VB=False):
        """Hard scatter of ``p0`` through ``process`` -> [Particle, Particle] or None below threshold (shower.py:467-507)."""
        ids = p0.get_ids()
        E0 = p0.get_pf()[0]
        if E0 <= np.max([self._minimum_calculable_energy[ids["PID"]], self.min_energy, ids["mass"]]):
            return None
        RM = np.array(p0.rotation_matrix(), dtype=float)
        x = self.draw_sample(E0, process=process, VB=VB)
        fid = self._next_shower_id
        self._next_shower_id += 1
        key = self._probe(capi.PROBE_PHILOX, 0, [[self._seed & 0xFFFFFFFF, self._seed >> 32, fid & 0xFFFFFFFF, fid >> 32, 0, 0xA0]], 2)
        inp = np.zeros((1, 8))
        inp[0, 0], inp[0, 1], inp[0, 6] = E0, ids["mass"], key[0, 0]
        inp[0, 2:2 + dimensionalities[process]] = x[:dimensionalities[process]]
        v = self._probe(capi.PROBE_KIN, process_code[process], inp, 8)[0]
        out = []
        for bit, pid in enumerate(process_PIDS[process]):
            pid = ids["PID"] if pid == 0 else pid
            four = v[4 * bit:4 * bit + 4]
            lab = np.concatenate([[four[0]], RM @ four[1:]])
            d = {"PID": pid, "parent_PID": ids["PID"], "ID": 2 * ids["ID"] + bit, "parent_ID": ids["ID"],
                 "generation_number": ids["generation_number"] + 1, "generation_process": process, "weight": ids["weight"],
                 "mass": mass_dict[pid]}
            out.append(Particle(lab, p0.get_rf(), d))
        return out

    def propagate_particle(self, Part0, Losses=False, MS=False):
        """Propagate one particle to its next hard scatter (shower.py:509-601): free path for photons, the dE/dx +
        multiple-scattering sub-step loop for e+-/mu+-.  Mutates ``Part0`` (pf, rf, ended) and returns it, like the reference.
        The draws are the engine's (Philox key from (seed, next shower id)); ``Losses`` must be False for photons and the
        target's dE/dx (what generate_shower passes, shower.py:624, 644) for charged particles."""
        if Part0.get_ended() is True:
            Part0.set_ended(True)
            return Part0
        ids = Part0.get_ids()
        pid = ids["PID"]
        charged = abs(pid) in (11, 13)
        if pid not in _STEPPING_PIDS:
            raise ValueError(f"PID {pid} is not propagated by the shower path")
        if charged and (Losses is False or abs(float(Losses) - self._dEdx * 0.1) > 1e-12 * self._dEdx):
            raise NotImplementedError("charged particles are propagated with the target's dE/dx (Losses = dEdx * 0.1), as generate_shower does")
        if not charged and Losses is not False:
            raise NotImplementedError("photons are propagated without energy loss (Losses=False), as generate_shower does")
        fid = self._next_shower_id
        self._next_shower_id += 1
        root = self._root_key(fid)
        inp = [[pid, *np.asarray(Part0.get_p0(), dtype=float), *np.asarray(Part0.get_r0(), dtype=float), ids["mass"], root[0], root[1], 1.0 if MS else 0.0]]
        o = self._probe(capi.PROBE_PROPAGATE, 0, inp, 9)[0]
        if o[8]:
            Part0.set_pf(np.array(o[0:4]))
            Part0.set_rf(np.array(o[4:7]))
        Part0.set_ended(True)
        return Part0

    def _root_key(self, shower_id):
        """Philox root key of a shower id (rng.cuh root_key): the first two output words of Philox(counter = id, key = seed)."""
        M0, M1, W0, W1, MASK = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85, 0xFFFFFFFF
        c = [shower_id & MASK, (shower_id >> 32) & MASK, 0, 0xA0]
        k0, k1 = self._seed & MASK, (self._seed >> 32) & MASK
        for _ in range(10):
            p0, p1 = M0 * c[0], M1 * c[2]
            c = [((p1 >> 32) ^ c[1] ^ k0) & MASK, p1 & MASK, ((p0 >> 32) ^ c[3] ^ k1) & MASK, p0 & MASK]
            k0, k1 = (k0 + W0) & MASK, (k1 + W1) & MASK
        return c[0], c[1]

    def _probe(self, what, process, inp, out_cols):
        inp = np.ascontiguousarray(inp, dtype=np.float64)
        out = np.zeros((len(inp), out_cols))
        capi.check(self._engine, capi.lib.pb_probe(self._engine, what, process, capi.dptr(inp), len(inp), inp.shape[1],
                                                   capi.dptr(out), out_cols))
        return out

    def replay(self, particles, tape, tape_off):
        """Replay mode (``pb_replay``): run ``n`` independent particle-steps - propagation, process choice, accept/reject
        sampling, kinematics - through the wave kernels' own device functions with every random number taken from a TAPE
        recorded from a reference run (layout: include/petite_b200.h).  ``particles`` (n, 10): pid, p0[4], r0[3], mass, flags
        (1 = multiple scattering on, 2 = short-lived, 4 = long-lived); ``tape`` (m,) doubles; ``tape_off`` (n + 1,) segment offsets.
        -> dict of per-step results (status 0 = the step consumed exactly its tape segment)."""
        part = np.ascontiguousarray(particles, dtype=np.float64).reshape(-1, 10)
        tape = np.ascontiguousarray(tape, dtype=np.float64)
        off = np.ascontiguousarray(tape_off, dtype=np.int64)
        n = len(part)
        assert len(off) == n + 1 and off[-1] <= len(tape)
        out = np.zeros((n, 32))
        capi.check(self._engine, capi.lib.pb_replay(self._engine, n, capi.dptr(part), capi.dptr(tape if len(tape) else np.zeros(1)),
                                                    off.ctypes.data_as(C.POINTER(C.c_int64)), capi.dptr(out)))
        return dict(status=out[:, 0].astype(int), nsub=out[:, 1].astype(int), process=out[:, 2].astype(int), ntrials=out[:, 3].astype(np.int64),
                    pf=out[:, 4:8], rf=out[:, 8:11], kept=out[:, 11].astype(int), pid_a=out[:, 12].astype(int), p_a=out[:, 13:17],
                    pid_b=out[:, 17].astype(int), p_b=out[:, 18:22], x=out[:, 22:26], consumed=out[:, 26].astype(np.int64),
                    weight_factor=out[:, 27], propagated=out[:, 28].astype(int), weight_factor_b=out[:, 29])

    def find_max(self, process, n_trials=100, seed=20261017, mT=None):
        """GPU ``do_find_max_work`` (utilities/find_maxes.py:55-119) for this target: -> (max_F (nE,), sigma (nE,))."""
        code = self._ALL_CODES[process]
        ms = (self._maps.get(process) if process in self._maps else getattr(self, "_dark_maps", {}).get(process))
        if ms is None:
            raise Exception("Process String does not match library")
        mf = np.zeros(len(ms.E)); sg = np.zeros(len(ms.E))
        if mT is None:
            mT = K.target_information[self._target_material]["mT"]
        capi.check(self._engine, capi.lib.pb_find_max(self._engine, code, int(n_trials), int(seed), float(mT), capi.dptr(mf), capi.dptr(sg)))
        return mf, sg

    def retrain_maps(self, processes=("Brem", "PairProd"), power=None, n_trials=100, seed=20261017, schedule=None):
        """Rows f-2 + f-1 on this engine: retrain the VEGAS maps of ``processes`` on the GPU at the energies of the loaded set
        (``petite_b200.train.Trainer``, training weight |jac f|^power, default 8: the accept/reject figure of merit), rebuild their
        ``max_F`` for this target (``find_max``, utilities/find_maxes.py:55-119) and switch the sampler to them.  The sampled
        distributions do not depend on the map (accept/reject is exact for any map whose ``max_F`` bounds jac f); the number of
        trials per sample does: 10 GeV photons in lead need ~14 with the shipped maps.  -> {process: (sigma / shipped-map sigma,
        efficiency / shipped-map efficiency)} per energy row."""
        from .train import Trainer, TRAIN_POWER
        tr = Trainer(device=self._device)
        out = {}
        for P in processes:
            old = self._maps[P]
            mf0, sg0 = self.find_max(P, n_trials=n_trials, seed=seed)
            grids, ninc, _ = tr.train(P, old.E, power=TRAIN_POWER if power is None else power, schedule=schedule, seed=seed)
            ms = tb.MapSet(P, old.E, ninc, grids, np.ones(len(old.E)), old.neval, old.Eg_min, old.Ee_min)
            self._upload_maps(process_code[P], ms)
            self._maps[P] = ms
            mf, sg = self.find_max(P, n_trials=n_trials, seed=seed)
            ms.max_F = mf
            self._upload_maps(process_code[P], ms)
            with np.errstate(all="ignore"):
                out[P] = (sg / sg0, (sg / mf) / (sg0 / mf0))
        return out

    def batch_from_particles(self, plist):
        """Upload an existing list of SM ``Particle`` objects as stack records (one pseudo-shower; fresh Philox keys)."""
        torch = self._torch
        n = len(plist)
        self._ensure_stack(max(n, 1024))
        p0 = np.array([np.asarray(p.get_p0(), dtype=float) for p in plist]).reshape(n, 4)
        pf = np.array([np.asarray(p.get_pf(), dtype=float) for p in plist]).reshape(n, 4)
        r0w = np.column_stack([np.array([np.asarray(p.get_r0(), dtype=float) for p in plist]).reshape(n, 3),
                               [p.get_ids()["weight"] for p in plist]])
        rf = np.column_stack([np.array([np.asarray(p.get_rf(), dtype=float) for p in plist]).reshape(n, 3),
                              [p.get_ids()["mass"] for p in plist]])
        gen = np.array([p.get_ids()["generation_number"] for p in plist], dtype=np.int64) & 0xFFFF
        meta = np.column_stack([[p.get_ids()["PID"] for p in plist], np.full(n, -1), (gen << 16) | 15, np.zeros(n)]).astype(np.int32)
        rng = np.random.default_rng((self._seed, self._next_shower_id))
        self._next_shower_id += 1
        key = rng.integers(0, 2 ** 32, size=(n, 2), dtype=np.uint64).astype(np.uint32).view(np.int32)
        t = self._stack_tensors
        for name, arr in (("p0", p0), ("pf", pf), ("r0w", r0w), ("rf", rf), ("meta", meta), ("key", key)):
            t[name][:n].copy_(torch.from_numpy(np.ascontiguousarray(arr)))
        t["ids"][:n, 6:8].copy_(torch.from_numpy(np.ascontiguousarray(r0w[:, 3]).view(np.int32).reshape(n, 2)))   # packed weight
        t["aux"][:n].zero_()
        b = ShowerBatch(self, t, n, {}, 1, 0)
        b.reference_order = lambda: (np.arange(n), np.array([0, n]))      # the list IS the reference order
        return b

    # ------------------------------------------------------------------ public stepping API
    def generate_showers(self, primaries, GlobalMS=True, capacity=None, first_shower_id=None):
        """Step many independent primaries (list of :class:`Particle`) at once -> :class:`ShowerBatch`."""
        return self.run_arrays(*self._pack_primaries(primaries), GlobalMS=GlobalMS, capacity=capacity,
                               first_shower_id=first_shower_id)

    def generate_showers_split(self, primaries, parts=2, GlobalMS=True, capacity=None, first_shower_id=None):
        """``generate_showers`` as ``parts`` concurrent sub-batches (:meth:`run_arrays_split`) -> list of :class:`ShowerBatch`."""
        return self.run_arrays_split(*self._pack_primaries(primaries), parts=parts, GlobalMS=GlobalMS, capacity=capacity,
                                     first_shower_id=first_shower_id)

    def generate_shower(self, p0, VB=False, GlobalMS=True):
        """One primary -> list of all particles of its shower, primary first (shower.py:603-708)."""
        if VB:
            print("Starting shower, initial particle with ID Info")
            print(p0.get_ids())
            print("Initial four-momenta:")
            print(p0.get_p0())
        p0.set_ended(False)
        if p0.get_p0()[0] < self.min_energy:
            p0.set_ended(True)
            return [p0.copy()]
        batch = self.generate_showers([p0], GlobalMS=GlobalMS)
        return batch.to_particles([p0])[0]
