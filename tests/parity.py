"""GPU-vs-oracle comparison helpers shared by tests/, smoke() and bench.py (checker only)."""
import numpy as np

from oracle.shower import OracleShower, OParticle
from oracle import consts as OC


def oracle_showers(prims, material, min_energy, seed, first_shower_id=0, **kw):
    o = OracleShower(None, material, min_energy, seed=seed, rng="counter", **kw)
    out = []
    for i, p in enumerate(prims):
        ids = p.get_ids()
        op = OParticle(p.get_p0(), p.get_r0(), PID=ids["PID"], ID=ids["ID"], gen=ids["generation_number"],
                       weight=ids["weight"], mass=ids["mass"], stability=ids["stability"])
        out.append(o.generate_shower(op, shower_id=first_shower_id + i))
    return out


def _relvec(a, b):
    a = np.asarray(a, dtype=float); b = np.asarray(b, dtype=float)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(a)), 1e-300))


def compare_with_oracle(batch, prims, material, min_energy, seed, **kw):
    """Shower by shower, in the reference's creation order: identical multiplicity, PIDs, processes, trial counts
    and sub-step counts ("decisions"); momenta/positions compared relative to the vector's largest component."""
    ref = oracle_showers(prims, material, min_energy, seed, first_shower_id=batch.first_shower_id, **kw)
    h = batch.to_host()
    order, offs = batch.reference_order()
    rep = dict(showers=len(prims), structure_mismatch=0, particles=0, max_rel_p0=0.0, max_rel_pf=0.0, max_abs_rf=0.0,
               max_rel_weight=0.0, first_mismatch=None)
    for i, olist in enumerate(ref):
        sl = order[offs[i]:offs[i + 1]]
        ok = len(sl) == len(olist)
        if ok:
            pid_o = np.array([q.PID for q in olist]); pr_o = np.array([OC.PROC_CODE[q.process] for q in olist])
            nt_o = np.array([q.ntrials for q in olist]); ns_o = np.array([q.nsub for q in olist])
            ok = (np.array_equal(h["pid"][sl], pid_o) and np.array_equal(h["process"][sl], pr_o)
                  and np.array_equal(h["ntrials"][sl], nt_o) and np.array_equal(h["nsub"][sl], ns_o))
        if not ok:
            rep["structure_mismatch"] += 1
            if rep["first_mismatch"] is None:
                rep["first_mismatch"] = dict(shower=i, n_gpu=len(sl), n_oracle=len(olist))
            continue
        rep["particles"] += len(sl)
        for s, q in zip(sl, olist):
            rep["max_rel_p0"] = max(rep["max_rel_p0"], _relvec(q.p0, h["p0"][s]))
            rep["max_rel_pf"] = max(rep["max_rel_pf"], _relvec(q.pf, h["pf"][s]))
            rep["max_rel_weight"] = max(rep["max_rel_weight"], abs(float(h["weight"][s]) - q.weight) / max(abs(q.weight), 1e-300))
            # positions relative to their own scale (at least 1 m): a track with n*sigma = 0 gets the reference's 1e12 m
            # mean free path (SURVEY Q-15) and ends ~1e10 m away
            rf = np.asarray(q.rf)
            rep["max_abs_rf"] = max(rep["max_abs_rf"], float(np.max(np.abs(rf - h["rf"][s])) / max(1.0, np.max(np.abs(rf)))))
    return rep
