"""Random-draw sources for the oracle.  TEST INFRASTRUCTURE - see oracle/__init__.py.

``StreamDraws``  consumes NumPy's legacy global stream and Python's ``random`` module in exactly the
                 order the reference does - used to pin the oracle against the
                 reference itself (tests/golden).
``CounterDraws`` derives every uniform from Philox(key = particle key, counter = (index, stream, ...))
                 (oracle/philox.py) - order-free, and what the CUDA engine implements.
"""
import random

import numpy as np

from . import philox as ph
from .physics import normals_from_uniforms


class StreamDraws:
    counter_mode = False

    def child(self, bit):
        return self

    def substep(self, i):
        return np.random.random(), np.random.random()

    def final(self):
        return np.random.random()

    def mcs(self, i, pc=0):
        sign = random.choice([-1, 1])
        z1 = random.gauss(0.0, 1.0)
        z2 = random.gauss(0.0, 1.0)
        return sign, z1, z2, random.random()

    def choice(self, pc=0):
        return np.random.random()

    def vegas_y(self, sweep, B, dim, pc):
        return np.random.random((B, dim))

    def vegas_u(self, sweep, B, dim, pc):
        return None  # drawn one by one with accept_u()

    def accept_u(self):
        return np.random.random()

    def kin(self, pc):
        return _Lazy2()

    def decay(self, pc=12):
        return np.random.random(), np.random.random()

    def decay_x(self, loop, i):
        # particle.py:371-372: np.random.uniform(0, x_max), np.random.uniform(0, 1) per accept/reject iteration
        return np.random.random(), np.random.random()

    def dbin(self, pc):
        return np.random.random()

    def pe(self, i, pc):
        return np.random.random(), np.random.random()

    def c0(self, pc):
        return np.random.random()


class _Lazy2:
    """Kinematics draw 0, 1 or 2 azimuths depending on the process: draw lazily, in order."""

    def __getitem__(self, i):
        return np.random.random()


class CounterDraws:
    counter_mode = True

    def __init__(self, key):
        self.key = key

    def child(self, bit):
        return CounterDraws(ph.child_key(self.key, bit))

    def _d(self, c0, st, c2=0, c3=0):
        a, b = ph.draw2(self.key, c0, st, c2, c3)
        return float(a), float(b)

    def substep(self, i):
        return self._d(i, ph.ST_SUBSTEP)

    def final(self):
        return self._d(0, ph.ST_FINAL)[0]

    def mcs(self, i, pc=0):
        # one call: (u_phi, u_radius) + the sign from a spare bit; only |z| = sqrt(z1^2 + z2^2) enters the reference's
        # angle (moliere.py:265-284), so the Box-Muller angle is fixed to 0 (z1 = |z|, z2 = 0)
        up, ur, spare = ph.draw2s(self.key, i, ph.ST_MCS, 0, pc)
        z1, z2 = normals_from_uniforms(0.0, float(ur))
        return (1 if int(spare) & 1 else -1), z1, z2, float(up)

    def choice(self, pc=0):
        return self._d(0, ph.ST_CHOICE, 0, pc)[0]

    def _trial_doubles(self, sweep, B, dim, pc):
        t = np.arange(sweep * B, (sweep + 1) * B, dtype=np.uint64)
        if dim == 4:        # two calls: y_0..y_3, and u_accept from the calls' spare bits
            a0, b0, s0 = ph.draw2s(self.key, t, ph.ST_VEGAS, 0, pc)
            a1, b1, s1 = ph.draw2s(self.key, t, ph.ST_VEGAS, 1, pc)
            return np.stack([a0, b0, a1, b1, ph.u48(s0, s1)], axis=1)
        ncall = (dim + 2) // 2
        cols = []
        for j in range(ncall):
            a, b = ph.draw2(self.key, t, ph.ST_VEGAS, j, pc)
            cols += [a, b]
        return np.stack(cols, axis=1)

    def vegas_y(self, sweep, B, dim, pc):
        self._last = self._trial_doubles(sweep, B, dim, pc)
        return self._last[:, :dim]

    def vegas_u(self, sweep, B, dim, pc):
        return self._last[:, dim]

    def kin(self, pc):
        return self._d(0, ph.ST_KIN, 0, pc)

    def decay(self, pc=12):
        return self._d(0, ph.ST_DECAY, 0, pc)

    def decay_x(self, loop, i):
        # i-th (x, u) pair of accept/reject loop 1 (path length), 2, 3 (the daughters' weights) of a decay in flight
        return self._d(i, ph.ST_DECAY, loop, 12)

    def dbin(self, pc):
        return self._d(0, ph.ST_DBIN, 0, pc)[0]

    def pe(self, i, pc):
        return self._d(i, ph.ST_PE, 0, pc)

    def c0(self, pc):
        return self._d(0, ph.ST_C0, 0, pc)[0]
