"""Host-side ``Particle`` with the reference's public interface (src/PETITE/particle.py:50-185).

On the GPU a particle is a record of the structure-of-arrays stack (include/petite_b200.h, ``pb_stack``); this
class is the per-object view the reference's users work with: the same constructor, the same id dictionary keys
and defaults (particle.py:22-28), the same getters/setters.  Only what the shower path needs is implemented
(three-body and long-lived decays are outside the hot-path scope).
"""
import numpy as np

from . import constants as K

mass_dict = dict(K.MASS)
meson_decay_dict = {k: [list(v) for v in vs] for k, vs in K.MESON_DECAYS.items()}
meson_twobody_branchingratios = {pid: opts[0][0] for pid, opts in meson_decay_dict.items()}

default_ids = {"PID": 11, "ID": 1, "parent_PID": 22, "parent_ID": -1, "generation_number": 0,
               "generation_process": "Input", "weight": 1.0, "mass": None, "stability": "stable",
               "production_time": 0.0, "decay_time": 0.0, "interaction_time": 0.0}


class Particle:
    """Container for one particle: creation state (p0, r0), state after propagation (pf, rf), ids."""

    def __init__(self, p0, r0=np.array([0, 0, 0]), id_dictionary=None):
        self._IDs = dict(default_ids)
        if id_dictionary:
            self._IDs.update({k: v for k, v in id_dictionary.items() if k in default_ids})
        self._mass = self._IDs["mass"]
        if isinstance(p0, list):
            p0 = np.array(p0)
        elif type(p0) in (int, float):
            # a bare number is an energy; momentum along +z (particle.py:83-86)
            if self._mass is None:
                self._mass = mass_dict[self._IDs["PID"]]
            p0 = np.array([p0, 0, 0, np.sqrt(p0 ** 2 - self._mass ** 2)])
        self.set_p0(p0)
        self.set_r0(np.array(r0) if isinstance(r0, list) else r0)
        self._Ended = False
        self._pf = p0
        self._rf = self._r0

    # ---- ids
    def set_ids(self, value):
        self._IDs = dict(default_ids)
        self._IDs.update({k: v for k, v in value.items() if k in default_ids})

    def get_ids(self):
        return self._IDs

    def update_ids(self, key, value):
        self._IDs[key] = value

    def get_pid(self):
        return self._IDs["PID"]

    def get_parent_pid(self):
        return self._IDs["parent_PID"]

    def get_weight(self):
        return self._IDs["weight"]

    def set_mass(self, value):
        self._mass = value

    # ---- kinematics
    def set_p0(self, value):
        self._p0 = value
        if self._mass is None:
            # mass back-computed and rounded to 6 decimals (particle.py:125-131; SURVEY Q-7)
            m = round(np.sqrt(round(value[0] ** 2 - value[1] ** 2 - value[2] ** 2 - value[3] ** 2, 12)), 6)
            self._mass = m
            self._IDs["mass"] = m

    def get_p0(self):
        return self._p0

    def set_pf(self, value):
        self._pf = value

    def get_pf(self):
        return self._pf

    def set_r0(self, value):
        self._r0 = value

    def get_r0(self):
        return self._r0

    def set_rf(self, value):
        self._rf = value

    def get_rf(self):
        return self._rf

    def get_angle_to_z_0(self):
        E0, px0, py0, pz0 = self.get_p0()
        return np.arccos(pz0 / np.sqrt(px0 ** 2 + py0 ** 2 + pz0 ** 2))

    def set_ended(self, value):
        if value != True and value != False:  # noqa: E712  (same acceptance rule as the reference)
            raise ValueError("Ended property must be a boolean.")
        self._Ended = value

    def get_ended(self):
        return self._Ended

    def copy(self):
        return Particle(self.get_p0(), self.get_r0(), self.get_ids())

    def lose_energy(self, value):
        """E -> max(E - value, m), direction kept (particle.py:143-153)."""
        E0, px0, py0, pz0 = self.get_pf()
        m = self._IDs["mass"]
        p30 = np.linalg.norm([px0, py0, pz0])
        E1 = max(E0 - value, m)
        p3f = np.sqrt(E1 ** 2 - m ** 2)
        if p3f > 0.0:
            self.set_pf([E1, px0 / p30 * p3f, py0 / p30 * p3f, pz0 / p30 * p3f])
        elif p3f == 0.0:
            self.set_pf([m, 0.0, 0.0, 0.0])

    def rotation_matrix(self):
        """Rz(phi) Ry(theta): z-hat -> direction of pf (particle.py:176-185)."""
        _, px, py, pz = self.get_pf()
        th = np.arccos(pz / np.sqrt(px ** 2 + py ** 2 + pz ** 2))
        ph = np.arctan2(py, px)
        ct, st, cp, sp = np.cos(th), np.sin(th), np.cos(ph), np.sin(ph)
        return [[ct * cp, -sp, st * cp], [ct * sp, cp, st * sp], [-st, 0, ct]]
