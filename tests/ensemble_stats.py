"""Per-shower summary statistics shared by the oracle fixture (tests/golden/make_ensemble.py) and the GPU ensemble tests.

Showers are the independent units: every observable below is ONE number (or one short histogram) per shower, so that a
two-sample KS test, or a chi-square on per-shower histogram means, is valid.  Definitions (records = every particle the
reference's generate_shower returns, primary included):

  mult      number of records                                   n_gamma / n_eplus   records with PID 22 / -11
  E_gamma   summed creation energy of the photons               Emax_sec  largest creation energy among secondaries
  z_mean    creation-energy-weighted mean of the end point z    rT_mean   mean transverse distance of the end points
  theta_e   mean polar angle of the created e+-                 spec      photons per shower in 8 log-spaced energy bins
  dark:  n_V, yield (weight sum), lw_med (median log10 weight), EV_mean, EV_max
"""
import numpy as np

m_e = 510.998950e-6
CONFIGS = {
    # BASELINE config 2 (10 GeV photon -> lead) and config 1 (10 GeV e- -> graphite) at reduced oracle statistics
    "c2_gamma_lead": dict(material="lead", pid=22, E0=10.0, mass=0.0, E_min=0.010, seed=20261017, n_oracle=600, mV=None),
    "c1_e_graphite": dict(material="graphite", pid=11, E0=10.0, mass=m_e, E_min=0.010, seed=20261017, n_oracle=600, mV=None),
    # BASELINE config 3 (dark shower, 10 GeV e- -> graphite) at the physically intended lightest trained mass
    "c3_dark_graphite": dict(material="graphite", pid=11, E0=10.0, mass=m_e, E_min=0.010, seed=20261017, n_oracle=240, mV=0.003),
    # BASELINE config 5, SM part (100 GeV mu- -> lead: MuonBrem / MuonE + multiple scattering); the oracle's dark-muon-brem pass
    # (~300 trials per emission, 6.6e3 emissions per shower) is too slow for a fixture and is covered by the replay tests instead
    "c5_mu_lead": dict(material="lead", pid=13, E0=100.0, mass=0.1056583755, E_min=0.010, seed=20261017, n_oracle=160, mV=None),
}
# Ensembles made by the UNMODIFIED reference in stream mode (tests/golden/make_ensemble_ref.py -> ensemble_ref.npz); config 5
# includes its dark pass (DarkMuonBrem first, as BASELINE.json lists the active processes)
REF_CONFIGS = {
    "c2_gamma_lead": dict(CONFIGS["c2_gamma_lead"], n_ref=500),
    "c1_e_graphite": dict(CONFIGS["c1_e_graphite"], n_ref=500),
    "c3_dark_graphite": dict(CONFIGS["c3_dark_graphite"], n_ref=200, active=["DarkBrem", "DarkAnn", "DarkComp"]),
    "c5_mu_lead_dark": dict(CONFIGS["c5_mu_lead"], n_ref=200, mV=0.030, active=["DarkMuonBrem", "DarkBrem", "DarkAnn", "DarkComp"]),
}
# BASELINE config 4: the 400 GeV table set (data_400GeV/, 4-D / 3-D maps retrained here: upstream lost them, tools/make_400GeV.py) and a
# synthetic beam-dump spectrum into lead, m_V = 10 MeV.  Primaries are i.i.d. draws (beam_dump_primaries), so a reference ensemble on
# the first n_ref of them and a GPU ensemble on 1e5 of them sample the same distribution.
REF_CONFIGS["c4_beamdump_lead_dark"] = dict(material="lead", pid=0, E0=400.0, mass=None, E_min=0.010, seed=20261017, n_ref=96, mV=0.010,
                                            active=["DarkBrem", "DarkAnn", "DarkComp"], data="data_400GeV")
SPEC_EDGES = np.logspace(-2, 1, 9)       # 8 bins, 10 MeV .. 10 GeV


def beam_dump_primaries(n, seed=20261017):
    """Config 4: 50 % photons resampled from the reference's 120 GeV pi0-photon beam (data_400GeV/Photons_From_Pi0s_120GeV.npy)
    scaled x (400 / 120) in momentum, 25 % e-, 25 % e+ with dN/dE ~ 1/E on [1, 400] GeV along +z; shuffled, so that any prefix is an
    i.i.d. sample of the spectrum.  -> (p (n,4), pid (n,), mass (n,))"""
    import os
    rng = np.random.default_rng(seed)
    beam = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "data_400GeV", "Photons_From_Pi0s_120GeV.npy"))
    kind = rng.integers(0, 4, n)                                  # 0, 1: photon; 2: e-; 3: e+
    g = beam[beam[:, 0] * (400.0 / 120.0) > 0.0016]
    pg = g[rng.integers(0, len(g), n)] * (400.0 / 120.0)
    E = np.exp(rng.uniform(np.log(1.0), np.log(400.0), n))
    pe = np.column_stack([E, np.zeros(n), np.zeros(n), np.sqrt(E ** 2 - m_e ** 2)])
    p = np.where((kind < 2)[:, None], pg, pe)
    pid = np.where(kind < 2, 22, np.where(kind == 2, 11, -11)).astype(np.int32)
    return p, pid, np.where(pid == 22, 0.0, m_e)


def config_primaries(cfg, n):
    """-> (p, pid, mass) arrays of a configuration's first n primaries (mono-energetic along +z, or the config-4 spectrum)."""
    if cfg["pid"] == 0:
        return beam_dump_primaries(n, cfg["seed"])
    E, m = cfg["E0"], cfg["mass"]
    return np.tile([E, 0.0, 0.0, np.sqrt(E * E - m * m)], (n, 1)), np.full(n, cfg["pid"], dtype=np.int32), np.full(n, m)


def summarise_oracle_sm(plist):
    E = np.array([q.p0[0] for q in plist]); pid = np.array([q.PID for q in plist])
    p0 = np.array([q.p0 for q in plist]); rf = np.array([q.rf for q in plist])
    sec = np.arange(len(plist)) > 0
    g = pid == 22
    el = (np.abs(pid) == 11) & sec
    th = np.arctan2(np.hypot(p0[:, 1], p0[:, 2]), p0[:, 3])
    return dict(mult=len(plist), n_gamma=int(g.sum()), n_eplus=int((pid == -11).sum()), E_gamma=float(E[g & sec].sum()),
                Emax_sec=float(E[sec].max()) if sec.any() else 0.0,
                z_mean=float(np.sum(E * rf[:, 2]) / np.sum(E)), rT_mean=float(np.mean(np.hypot(rf[:, 0], rf[:, 1]))),
                theta_e=float(th[el].mean()) if el.any() else 0.0,
                spec=np.histogram(E[g & sec], bins=SPEC_EDGES)[0].astype(np.float64))


def summarise_oracle_dark(vs):
    if not vs:
        return dict(n_V=0, dyield=0.0, lw_med=-300.0, EV_mean=0.0, EV_max=0.0)
    w = np.array([v.weight for v in vs]); E = np.array([v.p0[0] for v in vs])
    return dict(n_V=len(vs), dyield=float(w.sum()), lw_med=float(np.median(np.log10(np.maximum(w, 1e-300)))),
                EV_mean=float(E.mean()), EV_max=float(E.max()))


def summarise_gpu_sm(batch, n):
    """Same observables from the device-resident stack of a ShowerBatch, with torch reductions keyed by the shower id."""
    import torch
    t = batch._t
    m = batch.n
    sh = t["meta"][:m, 3].long()
    pid = t["meta"][:m, 0]
    parent = t["meta"][:m, 1]
    E = t["p0"][:m, 0]
    p0 = t["p0"][:m]
    rf = t["rf"][:m]
    sec = parent >= 0
    g = pid == 22
    f64 = lambda x: x.to(torch.float64)
    cnt = lambda mask: torch.bincount(sh[mask], minlength=n)
    wsum = lambda mask, w: torch.bincount(sh[mask], weights=w[mask], minlength=n)
    out = dict(mult=torch.bincount(sh, minlength=n), n_gamma=cnt(g), n_eplus=cnt(pid == -11), E_gamma=wsum(g & sec, E))
    emax = torch.zeros(n, dtype=torch.float64, device=E.device)
    emax.scatter_reduce_(0, sh[sec], E[sec], reduce="amax", include_self=True)
    out["Emax_sec"] = emax
    allm = torch.ones_like(sec)
    out["z_mean"] = wsum(allm, E * rf[:, 2]) / wsum(allm, E)
    out["rT_mean"] = wsum(allm, torch.hypot(rf[:, 0], rf[:, 1])) / f64(out["mult"])
    el = ((pid == 11) | (pid == -11)) & sec
    th = torch.atan2(torch.hypot(p0[:, 1], p0[:, 2]), p0[:, 3])
    ne = cnt(el)
    out["theta_e"] = torch.where(ne > 0, wsum(el, th) / f64(ne).clamp(min=1), torch.zeros_like(emax))
    gs = g & sec
    edges = torch.tensor(SPEC_EDGES, dtype=torch.float64, device=E.device)
    b = torch.bucketize(E[gs], edges, right=True) - 1          # np.histogram: [lo, hi) bins, last bin closed
    b = torch.where(E[gs] == edges[-1], torch.full_like(b, 7), b)
    ok = (b >= 0) & (b < 8)
    spec = torch.bincount(sh[gs][ok] * 8 + b[ok], minlength=8 * n).reshape(n, 8)
    out["spec"] = f64(spec)
    return {k: v.cpu().numpy() for k, v in out.items()}


def summarise_gpu_dark(dk, n):
    """Dark-vector observables from the device-resident dark stack (segmented median by two stable sorts)."""
    import torch
    t = dk._t
    m = dk.n
    sh = t["meta"][:m, 3].long()
    w = t["r0w"][:m, 3]
    E = t["p0"][:m, 0]
    n_V = torch.bincount(sh, minlength=n)
    dyield = torch.bincount(sh, weights=w, minlength=n)
    has = n_V > 0
    EV_mean = torch.where(has, torch.bincount(sh, weights=E, minlength=n) / n_V.clamp(min=1), torch.zeros_like(dyield))
    EV_max = torch.zeros(n, dtype=torch.float64, device=E.device)
    EV_max.scatter_reduce_(0, sh, E, reduce="amax", include_self=True)
    lw = torch.log10(w.clamp(min=1e-300))
    lw_sorted, o1 = torch.sort(lw)
    _, o2 = torch.sort(sh[o1], stable=True)
    lw_seg = lw_sorted[o2]                                  # grouped by shower, ascending inside each group
    start = torch.cumsum(n_V, 0) - n_V
    lo = (start + (n_V - 1).clamp(min=0) // 2).clamp(max=max(m - 1, 0))
    hi = (start + n_V // 2).clamp(max=max(m - 1, 0))
    med = 0.5 * (lw_seg[lo] + lw_seg[hi])                   # numpy's median: mean of the two middle values for even counts
    lw_med = torch.where(has, med, torch.full_like(med, -300.0))
    out = dict(n_V=n_V, dyield=dyield, lw_med=lw_med, EV_mean=EV_mean, EV_max=EV_max)
    return {k: v.cpu().numpy() for k, v in out.items()}
