"""Golden whole showers of LONG-LIVED primaries (pi+-, K+- decaying in flight: particle.py:363-389, 410-422)
made by the UNMODIFIED reference in stream mode (run in the build container, where /root/reference exists):

    python tests/golden/make_longlived.py        ->  tests/golden/longlived.npz

Same record layout as golden_showers() of make_golden.py; tests/test_oracle_golden.py replays the oracle on the same NumPy /
``random`` seeds and demands the same particles."""
import os
import random
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402  (imports the reference through _refstub)
from PETITE.particle import Particle, mass_dict  # noqa: E402
from PETITE.shower import Shower  # noqa: E402

CASES = [("graphite", 211, 20.0, 0.030, 301), ("lead", -211, 5.0, 0.030, 302), ("graphite", 321, 30.0, 0.030, 303),
         ("lead", -321, 8.0, 0.010, 304), ("graphite", -211, 2.0, 0.030, 305), ("lead", 211, 60.0, 0.030, 306)]


def main():
    mg.build_reference_dict_dir()
    proc_code = {"Input": 15, "SMDecay": 12}
    proc_code.update({p: i for i, p in enumerate(mg.SM_PROCESSES)})
    out = {}
    for k, (mat, pid, E, Emin, seed) in enumerate(CASES):
        s = Shower(mg.REFDIR, mat, Emin)
        m = mass_dict[pid]
        np.random.seed(seed)
        random.seed(seed)
        sh = s.generate_shower(Particle([E, 0, 0, np.sqrt(E ** 2 - m ** 2)], [0, 0, 0],
                                        {"PID": pid, "ID": 1, "mass": m, "stability": "long-lived"}))
        out[f"{k}/case"] = np.array([pid, E, Emin, seed, m])
        out[f"{k}/material"] = np.array(mat)
        out[f"{k}/pid"] = np.array([p.get_ids()["PID"] for p in sh])
        out[f"{k}/ID_mod"] = np.array([p.get_ids()["ID"] % (1 << 61) for p in sh], dtype=np.int64)
        out[f"{k}/gen"] = np.array([p.get_ids()["generation_number"] for p in sh])
        out[f"{k}/process"] = np.array([proc_code[p.get_ids()["generation_process"]] for p in sh])
        out[f"{k}/weight"] = np.array([p.get_ids()["weight"] for p in sh])
        out[f"{k}/mass"] = np.array([p.get_ids()["mass"] for p in sh], dtype=float)
        for name, get in (("p0", "get_p0"), ("pf", "get_pf"), ("r0", "get_r0"), ("rf", "get_rf")):
            out[f"{k}/{name}"] = np.array([getattr(p, get)() for p in sh], dtype=float)
        print("long-lived case", k, mat, pid, E, "->", len(sh), "particles; daughter weights", out[f"{k}/weight"][1:3], "decay point", out[f"{k}/rf"][0])
    out["n_cases"] = np.array(len(CASES))
    np.savez_compressed(os.path.join(HERE, "longlived.npz"), **out)


if __name__ == "__main__":
    main()
