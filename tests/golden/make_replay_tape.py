#!/usr/bin/env python
"""Record the random-number TAPE of whole reference showers, particle-step by particle-step, for pb_replay.

    python tests/golden/make_replay_tape.py        # -> tests/golden/replay_tape.npz   (build container only: needs /root/reference)

The UNMODIFIED reference (``Shower.generate_shower``, src/PETITE/shower.py:603-708) runs in stream mode while every random call it
makes - numpy.random.uniform / choice / random (shower.py:540, 561-562, 583, 671-697, 457; kinematics.py:33-325; the stub vegas
sweep of _refstub.py) and random.choice / gauss / uniform (moliere.py:284, 382) - is logged in consumption order.  The log is cut
into one segment per ``propagate_particle`` call, i.e. per particle-step, in the layout include/petite_b200.h documents for
pb_replay.  Stored next to it: the particle each step started from and what the reference made of it (pf, rf, chosen process,
number of accept/reject trials, the daughters it appended).  tests/test_gpu_replay.py feeds the tapes to the GPU decision code and
demands identical decisions and a fully consumed tape.
"""
import os
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
import make_golden  # noqa: E402
import PETITE.shower as rsh  # noqa: E402
from PETITE.particle import Particle  # noqa: E402
from PETITE.physical_constants import m_electron, m_muon  # noqa: E402
from oracle.consts import SM_PROCESSES  # noqa: E402

CASES = [("graphite", 11, 5.0, 0.010, 301), ("lead", 22, 5.0, 0.010, 302), ("lead", -11, 2.0, 0.010, 303), ("graphite", 13, 10.0, 0.030, 304),
         ("lead", 11, 1.0, 0.010, 305)]


class Recorder:
    """Wraps the generators' entry points; every wrapper returns bit for bit what the original would have returned."""

    def __init__(self):
        self.seg = None          # current particle-step's tape (list of floats)
        self.steps = []          # finished segments: dict(tape, particle, ...)
        self.block = None        # pending vegas sweep: (y rows, next row)
        self.o = dict(uniform=np.random.uniform, choice=np.random.choice, random=np.random.random, sample=np.random.random_sample,
                      pchoice=random.choice, gauss=random.gauss, puniform=random.uniform, prandom=random.random, draw_U=rsh.draw_U)

    # ---- numpy
    def np_uniform(self, low=0.0, high=1.0, size=None):
        assert size is None
        u = float(self.o["sample"]())
        v = low + (high - low) * u
        # kinematics azimuths (0, 2 pi) go on the tape as the uniform itself; U(6, 20) and U(0, 1) as drawn
        self.seg.append(u if (low == 0 and abs(high - 2.0 * np.pi) < 1e-12) else v)
        return v

    def np_choice(self, labels, p=None):
        u = float(self.o["sample"]())
        cdf = np.asarray(p, dtype=np.float64).cumsum()
        cdf /= cdf[-1]
        self.seg.append(u)
        lab = labels[int(cdf.searchsorted(u, side="right"))]
        self.cur["process"] = lab
        return lab

    def np_random(self, size=None):
        if size is None:                       # draw_U(): the accept uniform of the next untested point of the pending sweep
            u = float(self.o["sample"]())
            rows, k = self.block
            self.seg.extend(float(v) for v in rows[k])
            self.seg.append(u)
            self.block = (rows, k + 1)
            self.cur["ntrials"] += 1
            return u
        y = self.o["sample"](size)             # the stub vegas sweep: B x dim uniforms up front
        self.block = (y, 0)
        return y

    # ---- python random (moliere.py)
    def p_choice(self, seq):
        v = self.o["pchoice"](seq)
        self.seg.append(float(v))
        return v

    def p_gauss(self, mu, sigma):
        z = self.o["gauss"](0.0, 1.0)
        self.seg.append(z)
        return mu + z * sigma

    def p_uniform(self, a, b):
        u = self.o["prandom"]()
        self.seg.append(u)
        return a + (b - a) * u

    def install(self):
        np.random.uniform, np.random.choice, np.random.random = self.np_uniform, self.np_choice, self.np_random
        rsh.draw_U = self.np_random
        random.choice, random.gauss, random.uniform = self.p_choice, self.p_gauss, self.p_uniform

    def remove(self):
        np.random.uniform, np.random.choice, np.random.random = self.o["uniform"], self.o["choice"], self.o["random"]
        rsh.draw_U = self.o["draw_U"]
        random.choice, random.gauss, random.uniform = self.o["pchoice"], self.o["gauss"], self.o["puniform"]


def record_shower(material, pid, E, Emin, seed):
    s = rsh.Shower(make_golden.REFDIR, material, Emin)
    rec = Recorder()
    orig_prop = rsh.Shower.propagate_particle

    def prop(self, Part0, *a, **kw):
        ids = Part0.get_ids()
        rec.seg = []
        rec.cur = dict(tape=rec.seg, ID=ids["ID"], pid=ids["PID"], p0=np.array(Part0.get_p0(), dtype=float), r0=np.array(Part0.get_r0(), dtype=float),
                       mass=float(ids["mass"]), process="", ntrials=0)
        rec.steps.append(rec.cur)
        out = orig_prop(self, Part0, *a, **kw)
        rec.cur["pf"] = np.array(Part0.get_pf(), dtype=float)
        rec.cur["rf"] = np.array(Part0.get_rf(), dtype=float)
        return out
    m = {11: m_electron, -11: m_electron, 22: 0.0, 13: m_muon}[pid]
    np.random.seed(seed); random.seed(seed)
    rsh.Shower.propagate_particle = prop
    rec.install()
    try:
        sh = s.generate_shower(Particle([E, 0, 0, np.sqrt(E ** 2 - m ** 2)], [0, 0, 0], {"PID": pid, "ID": 1, "mass": m}))
    finally:
        rec.remove()
        rsh.Shower.propagate_particle = orig_prop
    # the reference run with the recorder in place must equal the plain run (the wrappers are transparent)
    np.random.seed(seed); random.seed(seed)
    plain = s.generate_shower(Particle([E, 0, 0, np.sqrt(E ** 2 - m ** 2)], [0, 0, 0], {"PID": pid, "ID": 1, "mass": m}))
    assert len(plain) == len(sh) and all(np.array_equal(np.asarray(a.get_p0(), float), np.asarray(b.get_p0(), float)) and
                                         np.array_equal(np.asarray(a.get_rf(), float), np.asarray(b.get_rf(), float)) for a, b in zip(plain, sh))
    by_id = {p.get_ids()["ID"]: p for p in sh}
    return rec.steps, by_id


def main():
    make_golden.build_reference_dict_dir()
    code = {p: i for i, p in enumerate(SM_PROCESSES)}
    out = {"n_cases": np.array(len(CASES))}
    for ci, (material, pid, E, Emin, seed) in enumerate(CASES):
        steps, by_id = record_shower(material, pid, E, Emin, seed)
        n = len(steps)
        part = np.zeros((n, 10)); res = np.zeros((n, 20)); off = np.zeros(n + 1, dtype=np.int64)
        tape = []
        for k, st in enumerate(steps):
            part[k] = [st["pid"], *st["p0"], *st["r0"], st["mass"], 1.0]
            tape.extend(st["tape"]); off[k + 1] = len(tape)
            d = [by_id.get(2 * st["ID"] + b) for b in (0, 1)]
            res[k, 0] = code.get(st["process"], -1)
            res[k, 1] = st["ntrials"]
            res[k, 2:6], res[k, 6:9] = st["pf"], st["rf"]
            res[k, 9] = (1 if d[0] is not None else 0) + (2 if d[1] is not None else 0)
            for b in (0, 1):
                if d[b] is not None:
                    res[k, 10 + 5 * b] = d[b].get_ids()["PID"]
                    res[k, 11 + 5 * b:15 + 5 * b] = np.asarray(d[b].get_p0(), dtype=float)
        pre = f"{ci}/"
        out[pre + "case"] = np.array([pid, E, Emin, seed]); out[pre + "material"] = np.array(material)
        out[pre + "particles"], out[pre + "tape"], out[pre + "tape_off"], out[pre + "ref"] = part, np.array(tape), off, res
        print("case", ci, material, pid, E, "steps", n, "particles", len(by_id), "tape doubles", len(tape))
    np.savez_compressed(os.path.join(HERE, "replay_tape.npz"), **out)
    print(os.path.getsize(os.path.join(HERE, "replay_tape.npz")), "bytes")


if __name__ == "__main__":
    main()
