"""Units (GeV), particle masses and target materials used by the host layer.

Numbers are those of the reference's ``physical_constants.py:11-58`` and ``particle.py:4-47``; the device copies
live in ``csrc/physics.cuh``.
"""
keV, MeV, GeV, TeV = 1e-6, 1e-3, 1.0, 1e3

alpha_em = 1.0 / 137.035999
m_electron = 510.998950 * keV
m_muon = 105.6583755 * MeV
m_proton = 938.272088 * MeV
m_proton_grams = 1.67262192369e-24
m_pi0 = 134.9768 * MeV
m_pi_pm = 139.57039 * MeV
m_K_pm = 493.677 * MeV
m_eta = 547.862 * MeV
m_eta_prime = 957.78 * MeV
m_omega = 782.65 * MeV
hbarc = 0.1973269804e-13          # GeV cm
GeVsqcm2 = hbarc ** 2
cmtom = 0.01


def _target(Z, A, mT, rho):
    return {"Z_T": Z, "A_T": A, "mT": mT, "rho": rho, "dEdx": 2.0 * rho}   # dEdx in MeV/cm


target_information = {
    "graphite": _target(6, 12, 11.178, 2.210),
    "lead": _target(82, 207, 207.2, 11.35),
    "iron": _target(26, 56, 55.845, 8.00),
    "hydrogen": _target(1, 1, 1.0, 1.0),
    "aluminum": _target(13, 27, 26.9815385, 2.699),
    "tungsten": _target(74, 183.84, 183.84, 19.3),
    "molybdenum": _target(42, 95.95, 95.95, 10.2),
}

MASS = {11: m_electron, -11: m_electron, 12: 0.0, -12: 0.0, 22: 0.0, 13: m_muon, -13: m_muon, 14: 0.0, -14: 0.0,
        111: m_pi0, 211: m_pi_pm, -211: m_pi_pm, 321: m_K_pm, -321: m_K_pm, 221: m_eta, 331: m_eta_prime,
        2212: m_proton, 223: m_omega}

MESON_DECAYS = {111: [[0.98823, [22, 22]]],
                221: [[0.3936, [22, 22]], [0.3257, [111, 111, 111]]],
                331: [[0.02307, [22, 22]], [0.224, [111, 111, 221]], [0.00250, [111, 111, 111]]],
                223: [[0.0828, [22, 111]]],
                211: [[0.9998, [-13, 14]]], -211: [[0.9998, [13, -14]]],
                321: [[0.6356, [-13, 14]]], -321: [[0.6356, [13, -14]]]}

# process codes shared with include/petite_b200.h (enum pb_process)
PROCESS_NAMES = ["Brem", "Ann", "PairProd", "Comp", "Moller", "Bhabha", "MuonE", "MuonBrem",
                 "DarkBrem", "DarkAnn", "DarkComp", "DarkMuonBrem", "SMDecay", "TwoBody_BSMDecay", "None", "Input"]
PROCESS_CODE = {n: i for i, n in enumerate(PROCESS_NAMES)}
