"""Units and constants (oracle copy).  TEST INFRASTRUCTURE - see oracle/__init__.py.

Values follow reference ``src/PETITE/physical_constants.py:11-58`` (GeV units).
"""
import math

TeV = 1e3
GeV = 1.0
MeV = 1e-3
keV = 1e-6

alpha_em = 1.0 / 137.035999                 # physical_constants.py:19
m_electron = 510.998950 * keV               # :20
m_proton = 938.272088 * MeV                 # :21
m_proton_grams = 1.67262192369e-24          # :22
n_avogadro = 6.0221409e+23                  # :25
hbarc = 0.1973269804e-13                    # :26  (GeV cm)
GeVsqcm2 = hbarc ** 2                       # :27
cmtom = 0.01                                # :28
m_muon = 105.6583755 * MeV                  # :30
m_pi0 = 134.9768 * MeV                      # :32
m_pi_pm = 139.57039 * MeV
m_K_pm = 493.677 * MeV
m_eta = 547.862 * MeV
m_eta_prime = 957.78 * MeV
m_omega = 782.65 * MeV

# physical_constants.py:49-58 ; dEdx in MeV/cm = 2*rho
TARGETS = {
    "graphite":   dict(Z_T=6,  A_T=12,     mT=11.178,     rho=2.210),
    "lead":       dict(Z_T=82, A_T=207,    mT=207.2,      rho=11.35),
    "iron":       dict(Z_T=26, A_T=56,     mT=55.845,     rho=8.00),
    "hydrogen":   dict(Z_T=1,  A_T=1,      mT=1.0,        rho=1.0),
    "aluminum":   dict(Z_T=13, A_T=27,     mT=26.9815385, rho=2.699),
    "tungsten":   dict(Z_T=74, A_T=183.84, mT=183.84,     rho=19.3),
    "molybdenum": dict(Z_T=42, A_T=95.95,  mT=95.95,      rho=10.2),
}
for _t in TARGETS.values():
    _t["dEdx"] = 2.0 * _t["rho"]

# particle.py:4-14
MASS = {11: m_electron, -11: m_electron, 12: 0.0, -12: 0.0, 22: 0.0,
        13: m_muon, -13: m_muon, 14: 0.0, -14: 0.0, 111: m_pi0,
        211: m_pi_pm, -211: m_pi_pm, 321: m_K_pm, -321: m_K_pm,
        221: m_eta, 331: m_eta_prime, 2212: m_proton, 223: m_omega}

# particle.py:15-20 (interaction lengths and c tau of the long-lived mesons, metres; physical_constants.py:31-33)
c_tau_pi_pm, c_tau_K_pm = 7.8045, 3.711
INT_LENGTH = {211: 1.796e-1, -211: 1.796e-1, 321: 2.2875e-1, -321: 2.2875e-1}
DECAY_LENGTH = {211: c_tau_pi_pm, -211: c_tau_pi_pm, 321: c_tau_K_pm, -321: c_tau_K_pm}

# particle.py:40-47 (first = two-body branching ratio)
MESON_DECAYS = {111: [[0.98823, [22, 22]]],
                221: [[0.3936, [22, 22]], [0.3257, [111, 111, 111]]],
                331: [[0.02307, [22, 22]], [0.224, [111, 111, 221]], [0.00250, [111, 111, 111]]],
                223: [[0.0828, [22, 111]]],
                211: [[0.9998, [-13, 14]]], -211: [[0.9998, [13, -14]]],
                321: [[0.6356, [-13, 14]]], -321: [[0.6356, [13, -14]]]}

# shower.py:36, dark_shower.py:49 ; the integer codes are this project's (shared with the CUDA engine)
SM_PROCESSES = ["Brem", "Ann", "PairProd", "Comp", "Moller", "Bhabha", "MuonE", "MuonBrem"]
DARK_PROCESSES = ["DarkBrem", "DarkAnn", "DarkComp", "DarkMuonBrem"]
PROC_CODE = {p: i for i, p in enumerate(SM_PROCESSES + DARK_PROCESSES)}
PROC_CODE["SMDecay"] = 12
PROC_CODE["TwoBody_BSMDecay"] = 13
PROC_CODE["Input"] = 15
PROC_DIM = {"Brem": 4, "Ann": 1, "PairProd": 4, "Comp": 1, "Moller": 1, "Bhabha": 1,
            "MuonE": 1, "MuonBrem": 4, "DarkBrem": 3, "DarkAnn": 1, "DarkComp": 1,
            "DarkMuonBrem": 3}
# shower.py:87-96 (0 -> parent PID)
PROC_PIDS = {"PairProd": [-11, 11], "Brem": [0, 22], "MuonBrem": [0, 22], "Comp": [11, 22],
             "Ann": [22, 22], "Moller": [0, 11], "Bhabha": [0, 11], "MuonE": [0, 11]}

TWO_PI = 2.0 * math.pi
