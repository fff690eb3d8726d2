// Device physics for the shower hot path: differential cross-sections in VEGAS-map variables, form factors,
// Lynch-Dahl multiple scattering, two-body kinematics.  fp64 throughout.  Each function cites the reference
// lines whose RESULTS it must reproduce (<=1e-12 relative away from cancellation points); the arithmetic is
// arranged for the GPU (shared sub-expressions, reciprocal reuse), not transliterated.
#pragma once
#include <math.h>
#include "rng.cuh"

namespace pb {

// physical_constants.py:19-30 (same literal expressions, so the doubles are bit-identical)
constexpr double kAlpha = 1.0 / 137.035999;
constexpr double kMe = 510.998950 * 1e-6;
constexpr double kMmu = 105.6583755 * 1e-3;
constexpr double kMp = 938.272088 * 1e-3;
constexpr double kMpi0 = 134.9768 * 1e-3;
constexpr double kMpipm = 139.57039 * 1e-3;   // physical_constants.py:32-33
constexpr double kMKpm = 493.677 * 1e-3;
constexpr double kPi = 3.141592653589793;
constexpr double kTwoPi = 2.0 * 3.141592653589793;
constexpr double kCmToM = 0.01;
constexpr double EGAMMA_MIN_KIN = 0.001;  // kinematics.py:9 (SURVEY Q-6)

enum Proc : int {
  P_BREM = 0, P_ANN = 1, P_PAIRPROD = 2, P_COMP = 3, P_MOLLER = 4, P_BHABHA = 5, P_MUONE = 6, P_MUONBREM = 7,
  P_DARKBREM = 8, P_DARKANN = 9, P_DARKCOMP = 10, P_DARKMUONBREM = 11, P_SMDECAY = 12, P_BSMDECAY = 13,
  P_NONE = 14, P_INPUT = 15, N_SAMPLED = 12
};

__host__ __device__ constexpr int proc_dim(int p) {
  return (p == P_BREM || p == P_PAIRPROD || p == P_MUONBREM) ? 4 : ((p == P_DARKBREM || p == P_DARKMUONBREM) ? 3 : 1);
}

// Per-engine constants derived on the host (glibc pow, as CPython uses) from Z, A, mV ... ; lives in __constant__.
struct Material {
  double Z, A, rho, dEdx;         // dEdx in GeV/m
  double mT;                      // sampler-time event_info['mT'] (= A, SURVEY Q-19)
  double min_energy, Eg_min, Ee_min;
  double fudge, rescale_mcs;
  double min_calc[5];             // e-, e+, gamma, mu-, mu+
  double ff_a0sq;                 // (184.15 * 2.718^-0.5 * Z^(-1/3) / m_e)^2        all_processes.py:92-103
  double ff_Z2a04;                // Z^2 * a0^4
  double dff_c1, dff_c2, dff_ap2; // all_processes.py:123-126
  double dff_ic2;                 // 1 / dff_c2
  double dff_inel_pref;           // Z / (c1^2 Z^2)
  double dff_pref;                // Z^2 c1^2
  double Z23;                     // Z^(2/3)   moliere.py:218
  double i2mT;                    // 1 / (2 mT)
  double me4, mV4;                // m_e**4, mV**4 as CPython computes them (libm pow)
  double mcs_C4, mcs_Cw, mcs_c3;  // Lynch-Dahl constants folded for mcs_fast
  long long max_trials;           // max_n_integrators * B
  // dark sector
  double mV, g_e, eps, Zeff, E_res_ann, E_thr_comp;
  int bound_electron, pad;
};

__device__ __forceinline__ int pid_class(int pid) {  // index into min_calc; -1 if not a stepping species
  switch (pid) { case 11: return 0; case -11: return 1; case 22: return 2; case 13: return 3; case -13: return 4; }
  return -1;
}
__host__ __device__ __forceinline__ double pid_mass(int pid) {  // particle.py:4-14 (mass_dict)
  switch (pid) {
    case 11: case -11: return kMe;
    case 13: case -13: return kMmu;
    case 111: return kMpi0;
    case 211: case -211: return kMpipm;
    case 321: case -321: return kMKpm;
    default: return 0.0;
  }
}

// numpy evaluates px**2 + py**2 + pz**2 left to right without contraction; keep the same roundings where the
// result feeds an ill-conditioned acos (particle.py:181).
__device__ __forceinline__ double norm3_nofma(double x, double y, double z) {
  return sqrt(__dadd_rn(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)), __dmul_rn(z, z)));
}

// Reciprocal without the IEEE division's special-case branch and slow-path call: MUFU.RCP64H seed (rel. error 2^-23) and two
// Newton steps (-> rounding level, <= 1 ulp).  Only for the re-associated hot-loop forms (ds_*_fast, mcs_fast, the sub-step),
// whose arguments are finite, normal and non-zero by construction and whose results are compared at 1e-12, not bit for bit.
__device__ __forceinline__ double fast_rcp(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  e = fma(-x, r, 1.0);
  return fma(r, e, r);
}

// Square root without the IEEE routine's special-case branch and slow-path call: MUFU.RSQ64H seed, one Newton step on 1/sqrt(x),
// one correction of x * r (<= 1 ulp).  Same contract as fast_rcp: finite normal positive arguments; anything else (negative, 0, Inf)
// yields NaN / garbage, which the branch-free integrands discard through their kinematic mask.
__device__ __forceinline__ double fast_sqrt(double x) {
  double r;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double h = 0.5 * x;
  r = fma(r, fma(-h * r, r, 0.5), r);          // r (1 + (1/2 - x r^2 / 2))
  r = fma(r, fma(-h * r, r, 0.5), r);
  double s = x * r;
  return fma(fma(-s, s, x), 0.5 * r, s);       // s + (x - s^2) r / 2
}

// fast_sqrt with sqrt(x <= 0) = 0 (the reference's sqrt gives exactly 0 at 0, e.g. for a particle that has just stopped)
__device__ __forceinline__ double fast_sqrt0(double x) { return x > 0.0 ? fast_sqrt(x) : 0.0; }
__device__ __forceinline__ double fast_rsqrt(double x) {       // 1 / sqrt(x), x finite, normal, positive; <= 1 ulp
  double r;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double h = 0.5 * x;
  r = fma(r, fma(-h * r, r, 0.5), r);
  return fma(r, fma(-h * r, r, 0.5), r);
}

// ---------------------------------------------------------------- hot-loop elementary functions
// The sub-step loop calls exp, log (twice), sincos and sincospi once per iteration.  libdevice's versions are accurate but
// ptxas materialises each of their ~60 fp64 polynomial coefficients as a pair of 32-bit immediate moves (24 % of k_loop's
// executed instructions, profiles/r01_summary.md).  The versions below keep the coefficients in constant memory, where a
// DFMA reads them as a c[bank][offset] operand, and drop the argument ranges the loop cannot produce.  Algorithms and
// coefficients are the classic fdlibm kernels (e_log.c, k_sin.c, k_cos.c; <= 1 ulp on the reduced range, checked against
// libm by tests/test_gpu_probes.py::test_hot_math).
__constant__ double kHotMath[] = {
  /* 0  Lg1..Lg7 */ 6.666666666666735130e-01, 3.999999999940941908e-01, 2.857142874366239149e-01, 2.222219843214978396e-01,
                    1.818357216161805012e-01, 1.531383769920937332e-01, 1.479819860511658591e-01,
  /* 7  ln2_hi, ln2_lo */ 6.93147180369123816490e-01, 1.90821492927058770002e-10,
  /* 9  S1..S6 */ -1.66666666666666324348e-01, 8.33333333332248946124e-03, -1.98412698298579493134e-04,
                  2.75573137070700676789e-06, -2.50507602534068634195e-08, 1.58969099521155010221e-10,
  /* 15 C1..C6 */ 4.16666666666666019037e-02, -1.38888888888741095749e-03, 2.48015872894767294178e-05,
                  -2.75573143513906633035e-07, 2.08757232129817482790e-09, -1.13596475577881948265e-11,
  /* 21 exp(-0.1) * (-1)^k / k!, k = 0..10 */
  0.9048374180359596, -0.9048374180359596, 0.4524187090179798, -0.1508062363393266,
  0.03770155908483165, -0.00754031181696633, 0.001256718636161055, -0.00017953123373729357,
  2.2441404217161696e-05, -2.4934893574624107e-06, 2.4934893574624105e-07,
  /* 32 pi_hi, pi_lo */ 3.141592653589793116e+00, 1.2246467991473532e-16,
  /* 34 log2(10), log10(2) hi (42 bits), lo, ln(10) */ 3.321928094887362, 0.30102999566395283, 2.8363394551044964e-14, 2.302585092994046,
  /* 38 1/k!, k = 2..13 */ 0.5, 0.16666666666666666, 0.041666666666666664, 0.008333333333333333, 0.001388888888888889,
  0.0001984126984126984, 2.48015873015873e-05, 2.7557319223985893e-06, 2.755731922398589e-07, 2.505210838544172e-08,
  2.08767569878681e-09, 1.6059043836821613e-10
};

// log(x) for finite normal x > 0 (e_log.c without the subnormal / special-value branches)
__device__ __forceinline__ double hot_log(double x) {
  const double* K = kHotMath;
  int hx = __double2hiint(x), lx = __double2loint(x);
  int k = (hx >> 20) - 1023;
  hx &= 0x000fffff;
  int i = (hx + 0x95f64) & 0x100000;
  double m = __hiloint2double(hx | (i ^ 0x3ff00000), lx);     // mantissa in [sqrt(1/2), sqrt(2))
  k += i >> 20;
  double f = m - 1.0;
  double s = f * fast_rcp(2.0 + f);
  double z = s * s, w = z * z;
  double t1 = w * fma(w, fma(w, K[5], K[3]), K[1]);
  double t2 = z * fma(w, fma(w, fma(w, K[6], K[4]), K[2]), K[0]);
  double R = t1 + t2;
  double hfsq = 0.5 * f * f;
  double dk = (double)k;
  return dk * K[7] - ((hfsq - fma(s, hfsq + R, dk * K[8])) - f);
}
// exp(-x) for x in [1/20, 1/6] (the hard-scatter test of a sub-step: x = 1 / U(6, 20)): degree-10 Taylor series about 0.1
__device__ __forceinline__ double hot_exp_neg_step(double x) {
  const double* K = kHotMath + 21;
  double d = x - 0.1;
  double r = K[10];
#pragma unroll
  for (int k = 9; k >= 0; --k) r = fma(r, d, K[k]);
  return r;
}
// 10^y for |y| < 300 (no overflow / subnormal handling): y = n log10(2) + r, 10^r = exp(r ln 10) by a degree-13 Taylor series
// (|r ln 10| <= ln(2)/2), scaled by 2^n through the exponent field.  <= 1 ulp against libm's pow(10, y).
__device__ __forceinline__ double hot_exp10(double y) {
  const double* K = kHotMath + 34;
  double n = rint(y * K[0]);
  double r = fma(-n, K[2], fma(-n, K[1], y));
  double z = r * K[3];
  double s = K[15];
#pragma unroll
  for (int k = 14; k >= 4; --k) s = fma(s, z, K[k]);
  s = fma(s, z, 1.0);
  s = fma(s, z, 1.0);
  return __hiloint2double(__double2hiint(s) + ((int)n << 20), __double2loint(s));
}
// sin and cos of x, |x| <= pi/4 (k_sin.c / k_cos.c)
__device__ __forceinline__ void hot_sincos_kernel(double x, double* sn, double* cs) {
  const double* K = kHotMath;
  double z = x * x;
  double rs = fma(z, fma(z, fma(z, fma(z, K[14], K[13]), K[12]), K[11]), K[10]);
  *sn = fma(z * x, fma(z, rs, K[9]), x);
  double rc = z * fma(z, fma(z, fma(z, fma(z, fma(z, K[20], K[19]), K[18]), K[17]), K[16]), K[15]);
  *cs = 1.0 - fma(0.5, z, -(z * rc));
}
__device__ __forceinline__ void hot_sincos(double x, double* sn, double* cs) {
  if (fabs(x) <= 0.78539816339744828) hot_sincos_kernel(x, sn, cs);
  else sincos(x, sn, cs);                                     // rare: multiple-scattering angles are small
}
// sin and cos of 2 pi u, u in [-1/2, 2): exact argument reduction (the reference rounds 2 pi u first: 4e-16 absolute apart)
__device__ __forceinline__ void hot_sincos_2pi(double u, double* sn, double* cs) {
  double t = 2.0 * u;
  double q = rint(2.0 * t);                                   // quadrant 0..4
  double r = fma(q, -0.5, t);                                 // exact, |r| <= 1/4
  double x = r * kHotMath[32];
  double xl = fma(r, kHotMath[32], -x) + r * kHotMath[33];    // low part of r * pi
  double s, c;
  hot_sincos_kernel(x, &s, &c);
  s = fma(xl, c, s);
  c = fma(-xl, s, c);
  int iq = (int)q & 3;
  double so = (iq & 1) ? c : s, co = (iq & 1) ? s : c;
  *sn = (iq & 2) ? -so : so;
  *cs = ((iq + 1) & 2) ? -co : co;
}

// cos(pi t), |t| <= 2: the azimuth factor of the two 4-D sampler integrands (once per accept/reject trial).  Exact reduction to
// |r| <= 1/4, both fdlibm kernels at r * pi (rounded once: <= 1.2e-16 absolute in the argument), quadrant select: <= 2 ulp, half the
// instructions of libdevice's cospi, which selects each of its coefficients with a pair of FSELs (9 % of k_sample, profiles/r02z).
__device__ __forceinline__ double hot_cospi(double t) {
  double a = fabs(t);
  double q = rint(2.0 * a);                                   // quadrant 0..4
  double r = fma(q, -0.5, a);                                 // exact
  double s, c;
  hot_sincos_kernel(r * kHotMath[32], &s, &c);
  int iq = (int)q;
  double v = (iq & 1) ? s : c;                                // cos(x + q pi/2): c, -s, -c, s, c
  return ((iq + 1) & 2) ? -v : v;
}
#ifndef PB_FAST_COSPI
#define PB_FAST_COSPI 1
#endif
#if PB_FAST_COSPI
#define PB_COSPI hot_cospi
#else
#define PB_COSPI cospi
#endif

// ---------------------------------------------------------------- form factors
__device__ __forceinline__ double ff_elastic(const Material& M, double t) {  // all_processes.py:99-103
  double den = 1.0 + M.ff_a0sq * t;
  return M.ff_Z2a04 * t * t / (den * den);
}

__device__ __forceinline__ double ff_el_inel_over_t2(const Material& M, double t) {  // all_processes.py:112-133
  const double mu_p = 2.79;
  double a = 1.0 / (1.0 + M.dff_c1 * t);
  double b = 1.0 + t / M.dff_c2;
  double Gel = a * a / (b * b);
  double r = M.dff_ap2 / (1.0 + M.dff_ap2 * t);
  double d = 1.0 + t / 0.71;
  double d2 = d * d;
  double Ginel = M.dff_inel_pref * (r * r) * ((1.0 + (mu_p * mu_p - 1.0) * t / (4.0 * kMp * kMp)) / (d2 * d2));
  return M.dff_pref * (Gel + Ginel);
}

// ---------------------------------------------------------------- differential cross-sections
// Each takes the sampled point x (map variables), the incoming energy and returns dsigma per unit map volume.

// all_processes.py:140-206 (dsigma_brem_dimensionless); ml = m_e (Brem) or m_mu (MuonBrem, SURVEY Q-5).
__device__ __forceinline__ double ds_brem(const Material& M, double ep, double ml, const double* x) {
  const double Egmin = M.Eg_min;
  double span = ep - ml - Egmin;
  double w = Egmin + x[0] * span;
  double k = ep / (2 * ml);
  double d = k * (x[1] + x[2]);
  double dp = k * (x[1] - x[2]);
  double ph = (x[3] - 0.5) * 2 * kPi;
  double epp = ep - w;
  bool ok = (Egmin < w) && (w < ep - ml) && (ml < epp) && (epp < ep) && (d > 0.0) && (dp > 0.0);
  double cph = cos(ph);
  double d2 = d * d, dp2 = dp * dp;
  double ml2 = ml * ml;
  double u = (1 + d2) / (2 * ep) - (1 + dp2) / (2 * epp);
  double qsq = ml2 * ((d2 + dp2 - 2 * d * dp * cph) + ml2 * u * u);
  double aem = kAlpha / ml;
  double PF = 8.0 / kPi * kAlpha * (aem * aem) * (epp * ml2 * ml2) / (w * ep * qsq * qsq) * d * dp;
  double jac = kPi * ep * ep * span / ml2;
  double FF = ff_elastic(M, qsq);
  double od = 1 + d2, odp = 1 + dp2;
  double T1 = d2 / (od * od);
  double T2 = dp2 / (odp * odp);
  double T3 = w * w / (2 * ep * epp) * (d2 + dp2) / (od * odp);
  double T4 = -(epp / ep + ep / epp) * (d * dp * cph) / (od * odp);
  // SURVEY Q-22: the reference multiplies by the boolean mask, so NaN*0 stays NaN (and is then rejected)
  return (ok ? 1.0 : 0.0) * PF * (T1 + T2 + T3 + T4) * jac * FF;
}

// all_processes.py:532-622 (dsigma_pairprod_dimensionless)
__device__ __forceinline__ double ds_pairprod(const Material& M, double w, const double* x) {
  const double me = kMe, me2 = kMe * kMe;
  double epp = me + x[0] * (w - 2 * me);
  double k = w / (2 * me);
  double dp = k * (x[1] + x[2]);
  double dm = k * (x[1] - x[2]);
  double ph = x[3] * 2 * kPi;
  double epm = w - epp;
  bool ok = (me < epm) && (epm < w) && (me < epp) && (epp < w) && (dm > 0.0) && (dp > 0.0);
  if (!ok) return 0.0;
  double cph = cos(ph);
  double dp2 = dp * dp, dm2 = dm * dm;
  double u = (1.0 + dp2) / (2.0 * epp) + (1.0 + dm2) / (2.0 * epm);
  double q2r = (dp2 + dm2 + 2.0 * dp * dm * cph) + me2 * u * u;
  double aem = kAlpha / me;
  double PF = 8.0 / kPi * kAlpha * (aem * aem) * epp * epm / (w * w * w * q2r * q2r) * dp * dm;
  double jac = kPi * w * w * (w - 2 * me) / me2;
  double FF = ff_elastic(M, me2 * q2r);
  double op = 1.0 + dp2, om = 1.0 + dm2;
  double T1 = -1.0 * dp2 / (op * op);
  double T2 = -1.0 * dm2 / (om * om);
  double T3 = w * w / (2.0 * epp * epm) * (dp2 + dm2) / (op * om);
  double T4 = (epp / epm + epm / epp) * (dp * dm * cph) / (op * om);
  return PF * (T1 + T2 + T3 + T4) * jac * FF;
}

// The 1-D integrands below contain differences of nearly equal terms (s + t at backward angles, 1 - b^2 ct^2 at high
// energy); they use only IEEE +,-,*,/,sqrt, so evaluating them in the reference's operation order WITHOUT fma
// contraction reproduces its doubles exactly.  m_ = mul, a_ = add, s_ = sub, d_ = div (round-to-nearest, never fused).
__device__ __forceinline__ double m_(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double a_(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double s_(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double d_(double a, double b) { return __ddiv_rn(a, b); }

// ---- folded forms of the two 4-D integrands for the sampling loop.  Same functions as ds_brem / ds_pairprod: the
// energy-only factors are computed once per sample (SampleConst), q^4 is cancelled analytically between the
// prefactor and the elastic form factor, and the ten divisions collapse to three.  Differences are re-association
// only (~1e-15 relative); points outside the kinematic mask return 0 (the reference returns 0 or NaN there - both
// are rejected by max_F*u < wgt*f).
struct SampleConst { double a, b, c, d, e; };

__device__ __forceinline__ SampleConst brem_const(const Material& M, double ep, double ml) {
  SampleConst s;
  s.a = ep - ml - M.Eg_min;                       // span
  s.b = ep / (2 * ml);                            // k
  s.c = 1.0 / ep;
  double aem = kAlpha / ml;
  double ml2 = ml * ml;
  s.d = (8.0 / kPi * kAlpha * (aem * aem)) * (ml2 * ml2) * s.c * (kPi * ep * ep * s.a / ml2) * M.ff_Z2a04;
  s.e = ml2;
  return s;
}
__device__ __forceinline__ double ds_brem_fast(const Material& M, const SampleConst& s, double ep, double ml, const double* x) {
  const double Egmin = M.Eg_min;
  double w = Egmin + x[0] * s.a;
  double d = s.b * (x[1] + x[2]);
  double dp = s.b * (x[1] - x[2]);
  double epp = ep - w;
  // branch-free: the kinematic mask selects at the end, so that the T trials a lane evaluates per round stay one basic block
  // (ILP); outside the mask the arithmetic below runs on garbage (possibly Inf / NaN) and is discarded
  bool ok = (Egmin < w) && (w < ep - ml) && (ml < epp) && (epp < ep) && (d > 0.0) && (dp > 0.0);
  double cph = PB_COSPI(2.0 * x[3] - 1.0);        // cos((x4 - 1/2) 2 pi)
  double d2 = d * d, dp2 = dp * dp;
  double od = 1 + d2, odp = 1 + dp2;
  double iepp = fast_rcp(epp);
  double u = od * (0.5 * s.c) - odp * (0.5 * iepp);
  double ddc = d * dp * cph;
  double qsq = s.e * ((d2 + dp2 - 2 * ddc) + s.e * u * u);
  double den = 1.0 + M.ff_a0sq * qsq;
  double io = fast_rcp(od * odp);
  double i1 = odp * io, i2 = od * io;             // 1/od, 1/odp
  double T = d2 * (i1 * i1) + dp2 * (i2 * i2) + (w * w) * (0.5 * s.c * iepp) * (d2 + dp2) * io - (epp * s.c + ep * iepp) * ddc * io;
  double f = s.d * (epp * d * dp) * fast_rcp(w * den * den) * T;
  return ok ? f : 0.0;
}

__device__ __forceinline__ SampleConst pairprod_const(const Material& M, double w) {
  const double me = kMe, me2 = kMe * kMe;
  SampleConst s;
  s.a = w - 2 * me;
  s.b = w / (2 * me);
  double aem = kAlpha / me;
  s.d = (8.0 / kPi * kAlpha * (aem * aem)) / (w * w * w) * (kPi * w * w * s.a / me2) * (M.ff_Z2a04 * (me2 * me2));
  s.c = M.ff_a0sq * me2;
  s.e = 0.0;
  return s;
}
__device__ __forceinline__ double ds_pairprod_fast(const SampleConst& s, double w, const double* x) {
  const double me = kMe, me2 = kMe * kMe;
  double epp = me + x[0] * s.a;
  double dp = s.b * (x[1] + x[2]);
  double dm = s.b * (x[1] - x[2]);
  double epm = w - epp;
  bool ok = (me < epm) && (epm < w) && (me < epp) && (epp < w) && (dm > 0.0) && (dp > 0.0);
  double cph = PB_COSPI(2.0 * x[3]);
  double dp2 = dp * dp, dm2 = dm * dm;
  double op = 1.0 + dp2, om = 1.0 + dm2;
  double ie = fast_rcp(epp * epm);
  double u = op * (0.5 * epm * ie) + om * (0.5 * epp * ie);
  double ddc = dp * dm * cph;
  double q2r = (dp2 + dm2 + 2.0 * ddc) + me2 * u * u;
  double den = 1.0 + s.c * q2r;
  double io = fast_rcp(op * om);
  double i1 = om * io, i2 = op * io;
  double T = -dp2 * (i1 * i1) - dm2 * (i2 * i2) + (w * w) * (0.5 * ie) * (dp2 + dm2) * io + (epp * epp + epm * epm) * ie * ddc * io;
  double f = s.d * (epp * epm * dp * dm) * fast_rcp(den * den) * T;
  return ok ? f : 0.0;
}

// all_processes.py:625-742 (dsigma_compton_dCT); mV > 0 is DarkComp.
__device__ __forceinline__ double ds_compton(const Material& M, double Eg, double mV, double ct) {
  const double me = kMe, me2 = kMe * kMe, me4 = M.me4;     // m**2 is an exact product; m**4 is libm pow (host)
  double s = a_(me2, m_(m_(2, Eg), me));
  double smv = a_(me, mV);
  if (s < m_(smv, smv)) return 0.0;
  double mV2 = m_(mV, mV);
  double smm = s_(s, mV2);
  double lam = sqrt(a_(s_(m_(smm, smm), m_(m_(2, me2), a_(s, mV2))), me4));
  double jac = m_(d_(s_(s, me2), m_(2, s)), lam);
  double mms = s_(mV2, s);
  double lam2 = sqrt(s_(a_(me4, m_(mms, mms)), m_(m_(2, me2), a_(mV2, s))));
  double ctl = m_(ct, lam2);
  double inner = s_(a_(me4, m_(s, a_(a_(-mV2, s), ctl))), m_(me2, a_(a_(mV2, m_(2, s)), ctl)));
  double t = d_(m_(-0.5, inner), s);
  double sm = s_(s, me2);
  double PF = d_(m_(m_(m_(2.0, kPi), kAlpha * kAlpha), 1.0), m_(sm, sm));
  double T1, T2, T3;
  if (mV == 0.0) {
    T1 = d_(s_(a_(m_(m_(6.0, me2), s), m_(3.0, me4)), m_(s, s)), m_(s_(me2, s), a_(a_(-me2, s), t)));
    double a = s_(a_(s, t), me2);
    T2 = d_(m_(4, me4), m_(a, a));
    double b = a_(s, me2);
    T3 = d_(a_(m_(t, sm), m_(b, b)), m_(sm, sm));
  } else {
    double mV4 = M.mV4;
    double a = s_(s_(a_(me2, mV2), s), t);
    double num = a_(a_(s_(s_(m_(m_(2.0, me2), s_(mV2, m_(3, s))), m_(3, me4)), m_(m_(2, mV2), s)), m_(2, mV4)), m_(s, s));
    T1 = d_(num, m_(s_(me2, s), a));
    T2 = d_(m_(m_(2, me2), a_(m_(2, me2), mV2)), m_(a, a));
    double c = s_(me2, s);
    T3 = d_(a_(m_(a_(me2, s), a_(a_(me2, mV2), s)), m_(t, sm)), m_(c, c));
  }
  return m_(m_(PF, jac), a_(a_(T1, T2), T3));
}

// all_processes.py:469-529 (dsigma_annihilation_dCT)
__device__ __forceinline__ double ds_annihilation(double Ee, double mV, double EgMin, double ct) {
  const double me = kMe;
  double s = m_(m_(2.0, me), a_(Ee, me));
  double mV2 = m_(mV, mV);
  double ctMax = d_(m_(sqrt(d_(a_(Ee, me), s_(Ee, me))), s_(m_(m_(2, me), a_(s_(Ee, m_(2, EgMin)), me)), mV2)),
                    s_(m_(m_(2, me), a_(Ee, me)), mV2));
  if (s < mV2) return 0.0;
  if (ct > ctMax) return 0.0;
  double b = sqrt(s_(1.0, d_(m_(4.0, m_(me, me)), s)));
  double ct2 = m_(ct, ct);
  double pre = d_(m_(m_(4.0, kPi), kAlpha * kAlpha), m_(s, s_(1, m_(m_(b, b), ct2))));
  double smv = s_(s, mV2);
  double br = a_(m_(d_(smv, m_(2, s)), a_(1, ct2)), d_(m_(2.0, mV2), smv));
  return m_(pre, br);
}

// all_processes.py:787-840 (dsigma_moller_dCT)
__device__ __forceinline__ double ds_moller(double Ee, double DE, double ct) {
  const double me = kMe, me2 = kMe * kMe;
  double lim = 2.0 * DE / (Ee - me);
  if (!((ct > -1 + lim) && (ct < 1.0 - lim))) return 0.0;
  double s = me2 + 2 * Ee * me;
  double c2 = ct * ct, c4 = c2 * c2;
  double num = s * s * (3 + c2) * (3 + c2) - 8 * me2 * s * (7 + c4) + 16 * me2 * me2 * (6 - 3 * c2 + c4);
  double a = s - 4 * me2;
  return 16 * kPi * kPi * kAlpha * kAlpha * num / (8 * kPi * s * a * a * (1 - ct) * (1 - ct) * (1 + ct) * (1 + ct));
}

// all_processes.py:948-1007 (dsigma_bhabha_dCT)
__device__ __forceinline__ double ds_bhabha(double Ee, double DE, double ct) {
  const double m = kMe;
  double lim = 2.0 * DE / (Ee - m);
  if (!((ct > -1 + lim) && (ct < 1.0 - lim))) return 0.0;
  double m2 = m * m, m4 = m2 * m2, m6 = m4 * m2, m8 = m4 * m4;
  double s = m2 + 2 * Ee * m;
  double s2 = s * s, s3 = s2 * s, s4 = s2 * s2;
  double c1 = -1 + ct;
  double num = 256 * c1 * c1 * ct * ct * m8
             - 128 * c1 * (1 + ct * (1 + ct) * (-3 + 2 * ct)) * m6 * s
             + 16 * (7 + ct * (2 + ct * (-5 + 6 * c1 * ct))) * m4 * s2
             - 8 * (7 + ct * (-3 + ct * (3 + ct * (-1 + 2 * ct)))) * m2 * s3
             + (3 + ct * ct) * (3 + ct * ct) * s4;
  double a = -4 * m2 + s;
  return (kAlpha * kAlpha * kPi * num) / (2 * c1 * c1 * s3 * a * a);
}

// all_processes.py:843-886 (dsigma_muonelectron_dCT)
__device__ __forceinline__ double ds_muone(double Emu, double DE, double ct) {
  const double me = kMe, me2 = kMe * kMe, mm2 = kMmu * kMmu;
  double s = me2 + mm2 + 2 * me * Emu;
  double t_limit = 2.0 * me * (me - DE);
  double a = s + me2 - mm2;
  double t = -2.0 * (1 - ct) * (a * a / (4.0 * s) - me2);
  if (!(t < t_limit)) return 0.0;
  double b = s + t - 4 * me2;
  return 16 * kPi * kPi * kAlpha * kAlpha * (s * s + 2 * (me2 + mm2) * (2 * t + mm2 - 3 * me2) + b * b) / (16 * kPi * s * t * t);
}

// all_processes.py:208-372 (dsig_dx_dcostheta_dark_brem_exact_tree_level, Method "Log")
__device__ __forceinline__ double ds_darkbrem(const Material& M, double Eb, double ml, const double* xx) {
  const double mV = M.mV, MT = M.mT;
  const double LN10 = 2.302585092994046;
  double x = xx[0];
  double omc = pow(10.0, xx[1]);
  double cth = 1.0 - omc;
  double ttilde = pow(10.0, xx[2]);
  double Jac = omc * ttilde * (LN10 * LN10);
  double xE = x * Eb;
  double mV2 = mV * mV, ml2 = ml * ml;
  double k = sqrt(fabs(xE * xE - mV2));
  double p = sqrt(Eb * Eb - ml2);
  double V2 = p * p + k * k - 2 * p * k * cth;
  double V = sqrt(V2);
  double utilde = -2 * (x * Eb * Eb - k * p * cth) + mV2;
  double Er = (1 - x) * Eb + MT;
  double discr = utilde * utilde + 4 * MT * utilde * Er + 4 * MT * MT * (V * V);
  double sq = sqrt(fabs(discr));
  double den = 2 * Er * Er - 2 * (V * V);
  double cmn = V * (utilde + 2 * MT * Er);
  double Qp = fabs((cmn + Er * sq) / den);
  double Qm = fabs((cmn - Er * sq) / den);
  double tplus = 2 * MT * (sqrt(MT * MT + Qp * Qp) - MT);
  double tminus = 2 * MT * (sqrt(MT * MT + Qm * Qm) - MT);
  double tc = 2 * MT * (MT + Eb) * sqrt(Eb * Eb + ml2) / (MT * (MT + 2 * Eb) + ml2);
  double tconv = tc * tc;
  double t = ttilde * tconv;
  double q0 = -t / (2 * MT);
  double q = sqrt(t * t / (4 * MT * MT) + t);
  double e3 = Eb + q0 - xE;
  double cthq = -((V * V) + q * q + ml2 - e3 * e3) / (2 * V * q);
  double mm = mV2 + 2 * ml2;
  double Y = -t + 2 * q0 * Eb - 2 * q * p * (p - k * cth) * cthq / V;
  double W = fabs(Y * Y - 4 * q * q * p * p * k * k * (1 - cth * cth) * (1 - cthq * cthq) / (V * V));
  bool ok = (xE >= mV) && (discr >= 0) && (tplus > tminus) && (t > tminus) && (t < tplus) && (fabs(cthq) <= 1.0) && (W > 0);
  if (!ok) return 0.0;
  double Am2 = -8 * MT * (4 * Eb * Eb * MT - t * (2 * Eb + MT)) * mm;
  double A1 = 8 * MT * MT / utilde;
  double Am1 = (8 / utilde) * (MT * MT * (2 * t * utilde + utilde * utilde
                                          + 4 * Eb * Eb * (2 * (x - 1) * mm - t * ((x - 2) * x + 2))
                                          + 2 * t * (-mV2 + 2 * ml2 + t))
                               - 2 * Eb * MT * t * ((1 - x) * utilde + (x - 2) * (mm + t))
                               + t * t * (utilde - mV2));
  double A0 = (8 / (utilde * utilde)) * (MT * MT * (2 * t * utilde + (t - 4 * Eb * Eb * (x - 1) * (x - 1)) * mm)
                                         + 2 * Eb * MT * t * (utilde - (x - 1) * mm));
  double sW = sqrt(W);
  double phi_int = (A0 + Y * A1 + Am1 / sW + Y * Am2 / (W * sW)) / (8 * MT * MT);
  double FF = ff_el_inel_over_t2(M, t);
  double ans = FF * (kAlpha * kAlpha * kAlpha) * k * Eb * phi_int / (p * sqrt(k * k + p * p - 2 * p * k * cth));
  return ans * tconv * Jac;
}

// ---- the dark-brem integrand as the sampler evaluates it: same formula and operation order in the cancellation-prone
// parts (utilde, discr, Y, W), but 10^x through hot_exp10, the energy-only factors hoisted per sample (SampleConst:
// b = |p|, c = tconv, d = 1 / (2 mT), e = 1 / |p|), divisions by shared denominators as branch-free reciprocals, and the
// kinematic cuts tested as soon as their inputs exist.  Agrees with ds_darkbrem to ~1e-13 relative away from W -> 0.
__device__ __forceinline__ SampleConst darkbrem_const(const Material& M, double Eb, double ml) {
  const double MT = M.mT, ml2 = ml * ml;
  SampleConst s;
  s.a = 0.0;
  s.b = sqrt(Eb * Eb - ml2);
  double tc = 2 * MT * (MT + Eb) * sqrt(Eb * Eb + ml2) / (MT * (MT + 2 * Eb) + ml2);
  s.c = tc * tc;
  s.d = 1.0 / (2 * MT);
  s.e = 1.0 / s.b;
  return s;
}
__device__ __forceinline__ double ff_el_inel_over_t2_fast(const Material& M, double t) {
  const double mu_p = 2.79;
  double ia = 1.0 + M.dff_c1 * t, b = 1.0 + t * M.dff_ic2;
  double g = fast_rcp(ia * b);
  double Gel = g * g;
  double r = M.dff_ap2 * fast_rcp(1.0 + M.dff_ap2 * t);
  double d = 1.0 + t * (1.0 / 0.71);
  double d2 = d * d;
  double Ginel = M.dff_inel_pref * (r * r) * ((1.0 + t * ((mu_p * mu_p - 1.0) / (4.0 * kMp * kMp))) * fast_rcp(d2 * d2));
  return M.dff_pref * (Gel + Ginel);
}
#ifndef PB_DB_BRANCHFREE
#define PB_DB_BRANCHFREE 0      // measured (profiles/r02_summary.md): with ~95 % of the dark-brem trials outside the kinematic cuts, a warp
#endif                          // whose lanes ALL leave early is common enough that the early returns win (config 3: 67.7 vs 85.1 ms)
#ifndef PB_DB_FASTSQRT
#define PB_DB_FASTSQRT 1
#endif
#if PB_DB_FASTSQRT
#define DB_SQRT fast_sqrt
#else
#define DB_SQRT sqrt
#endif
// the same integrand with the kinematic cuts tested as soon as their inputs exist (early returns)
__device__ __forceinline__ double ds_darkbrem_fast_eo(const Material& M, const SampleConst& sc, double Eb, double ml, const double* xx) {
  const double mV = M.mV, MT = M.mT;
  const double LN10 = 2.302585092994046;
  double x = xx[0];
  double xE = x * Eb;
  if (!(xE >= mV)) return 0.0;
  double omc = hot_exp10(xx[1]);
  double cth = 1.0 - omc;
  double ttilde = hot_exp10(xx[2]);
  double mV2 = mV * mV, ml2 = ml * ml;
  double k = DB_SQRT(fabs(xE * xE - mV2));
  const double p = sc.b;
  double V2 = p * p + k * k - 2 * p * k * cth;
  double V = DB_SQRT(V2);
  double VV = V * V;
  double utilde = -2 * (x * Eb * Eb - k * p * cth) + mV2;
  double Er = (1 - x) * Eb + MT;
  double discr = utilde * utilde + 4 * MT * utilde * Er + 4 * MT * MT * VV;
  if (!(discr >= 0)) return 0.0;
  double sq = DB_SQRT(discr);
  double iden = fast_rcp(2 * Er * Er - 2 * VV);
  double cmn = V * (utilde + 2 * MT * Er);
  double Qp = fabs((cmn + Er * sq) * iden);
  double Qm = fabs((cmn - Er * sq) * iden);
  double tplus = 2 * MT * (DB_SQRT(MT * MT + Qp * Qp) - MT);
  double tminus = 2 * MT * (DB_SQRT(MT * MT + Qm * Qm) - MT);
  const double tconv = sc.c, i2MT = sc.d;
  double t = ttilde * tconv;
  if (!((tplus > tminus) && (t > tminus) && (t < tplus))) return 0.0;
  double q0 = -t * i2MT;
  double q = DB_SQRT(t * t * (i2MT * i2MT) + t);
  double e3 = Eb + q0 - xE;
  double iV = fast_rcp(V), iq = fast_rcp(q);
  double cthq = -(VV + q * q + ml2 - e3 * e3) * (0.5 * iV * iq);
  double mm = mV2 + 2 * ml2;
  double Y = -t + 2 * q0 * Eb - 2 * q * p * (p - k * cth) * cthq * iV;
  double W = fabs(Y * Y - 4 * q * q * p * p * k * k * (1 - cth * cth) * (1 - cthq * cthq) * (iV * iV));
  if (!((fabs(cthq) <= 1.0) && (W > 0))) return 0.0;
  double iu = fast_rcp(utilde);
  double Am2 = -8 * MT * (4 * Eb * Eb * MT - t * (2 * Eb + MT)) * mm;
  double A1 = 8 * MT * MT * iu;
  double Am1 = (8 * iu) * (MT * MT * (2 * t * utilde + utilde * utilde
                                      + 4 * Eb * Eb * (2 * (x - 1) * mm - t * ((x - 2) * x + 2))
                                      + 2 * t * (-mV2 + 2 * ml2 + t))
                           - 2 * Eb * MT * t * ((1 - x) * utilde + (x - 2) * (mm + t))
                           + t * t * (utilde - mV2));
  double A0 = (8 * iu * iu) * (MT * MT * (2 * t * utilde + (t - 4 * Eb * Eb * (x - 1) * (x - 1)) * mm)
                               + 2 * Eb * MT * t * (utilde - (x - 1) * mm));
  double sW = DB_SQRT(W);
  double isW = fast_rcp(sW);
  double phi_int = (A0 + Y * A1 + Am1 * isW + Y * Am2 * (isW * isW * isW)) * (0.5 * i2MT * i2MT);
  double FF = ff_el_inel_over_t2_fast(M, t);
  double Jac = omc * ttilde * (LN10 * LN10);
  return FF * (kAlpha * kAlpha * kAlpha) * k * Eb * phi_int * (sc.e * iV) * tconv * Jac;
}

__device__ __forceinline__ double ds_darkbrem_fast_bf(const Material& M, const SampleConst& sc, double Eb, double ml, const double* xx) {
  // branch-free (the kinematic cuts select at the end): the lanes of a warp sit at independent points, so an early return saves
  // nothing under SIMT, and straight-line code lets the T trials a lane evaluates per round interleave.  Outside the cuts the
  // arithmetic runs on garbage (sqrt of a negative number -> NaN) and is discarded.
  const double mV = M.mV, MT = M.mT;
  const double LN10 = 2.302585092994046;
  double x = xx[0];
  double xE = x * Eb;
  double omc = hot_exp10(xx[1]);
  double cth = 1.0 - omc;
  double ttilde = hot_exp10(xx[2]);
  double mV2 = mV * mV, ml2 = ml * ml;
  double k = DB_SQRT(fabs(xE * xE - mV2));
  const double p = sc.b;
  double V2 = p * p + k * k - 2 * p * k * cth;
  double V = DB_SQRT(V2);
  double VV = V * V;
  double utilde = -2 * (x * Eb * Eb - k * p * cth) + mV2;
  double Er = (1 - x) * Eb + MT;
  double discr = utilde * utilde + 4 * MT * utilde * Er + 4 * MT * MT * VV;
  bool ok = (xE >= mV) && (discr >= 0);
  double sq = DB_SQRT(discr);
  double iden = fast_rcp(2 * Er * Er - 2 * VV);
  double cmn = V * (utilde + 2 * MT * Er);
  double Qp = fabs((cmn + Er * sq) * iden);
  double Qm = fabs((cmn - Er * sq) * iden);
  double tplus = 2 * MT * (DB_SQRT(MT * MT + Qp * Qp) - MT);
  double tminus = 2 * MT * (DB_SQRT(MT * MT + Qm * Qm) - MT);
  const double tconv = sc.c, i2MT = sc.d;
  double t = ttilde * tconv;
  ok = ok && (tplus > tminus) && (t > tminus) && (t < tplus);
  double q0 = -t * i2MT;
  double q = DB_SQRT(t * t * (i2MT * i2MT) + t);
  double e3 = Eb + q0 - xE;
  double iV = fast_rcp(V), iq = fast_rcp(q);
  double cthq = -(VV + q * q + ml2 - e3 * e3) * (0.5 * iV * iq);
  double mm = mV2 + 2 * ml2;
  double Y = -t + 2 * q0 * Eb - 2 * q * p * (p - k * cth) * cthq * iV;
  double W = fabs(Y * Y - 4 * q * q * p * p * k * k * (1 - cth * cth) * (1 - cthq * cthq) * (iV * iV));
  ok = ok && (fabs(cthq) <= 1.0) && (W > 0);
  double iu = fast_rcp(utilde);
  double Am2 = -8 * MT * (4 * Eb * Eb * MT - t * (2 * Eb + MT)) * mm;
  double A1 = 8 * MT * MT * iu;
  double Am1 = (8 * iu) * (MT * MT * (2 * t * utilde + utilde * utilde
                                      + 4 * Eb * Eb * (2 * (x - 1) * mm - t * ((x - 2) * x + 2))
                                      + 2 * t * (-mV2 + 2 * ml2 + t))
                           - 2 * Eb * MT * t * ((1 - x) * utilde + (x - 2) * (mm + t))
                           + t * t * (utilde - mV2));
  double A0 = (8 * iu * iu) * (MT * MT * (2 * t * utilde + (t - 4 * Eb * Eb * (x - 1) * (x - 1)) * mm)
                               + 2 * Eb * MT * t * (utilde - (x - 1) * mm));
  double sW = DB_SQRT(W);
  double isW = fast_rcp(sW);
  double phi_int = (A0 + Y * A1 + Am1 * isW + Y * Am2 * (isW * isW * isW)) * (0.5 * i2MT * i2MT);
  double FF = ff_el_inel_over_t2_fast(M, t);
  double Jac = omc * ttilde * (LN10 * LN10);
  double f = FF * (kAlpha * kAlpha * kAlpha) * k * Eb * phi_int * (sc.e * iV) * tconv * Jac;
  return ok ? f : 0.0;
}

__device__ __forceinline__ double ds_darkbrem_fast(const Material& M, const SampleConst& sc, double Eb, double ml, const double* xx) {
#if PB_DB_BRANCHFREE
  return ds_darkbrem_fast_bf(M, sc, Eb, ml, xx);
#else
  return ds_darkbrem_fast_eo(M, sc, Eb, ml, xx);
#endif
}

// radiative_return.py:26-79 + all_processes.py:400-466 (dsigma_radiative_return_du)
__device__ __forceinline__ double kf_beta(double s) { return (2.0 * kAlpha / kPi) * (log(s / (kMe * kMe)) - 1.0); }
__device__ __forceinline__ double fl_kf(double x, double beta) {
  if (x >= 1.0) x = 1.0 - 1e-10;
  return (beta / 16.0) * ((8.0 + 3.0 * beta) * pow(1.0 - x, beta / 2.0 - 1.0) - 4.0 * (1.0 + x));
}
__device__ __forceinline__ double fl_kf_scaled(double x, double beta) {
  return (beta / 16.0) * ((8.0 + 3.0 * beta) - 4.0 * (1.0 + x) * pow(1.0 - x, 1.0 - beta / 2.0));
}
__device__ __forceinline__ double ds_darkann(const Material& M, double Ee, double u0) {
  const double me = kMe;
  double mV2 = M.mV * M.mV;
  double s = 2.0 * me * (Ee + me);
  if (s < mV2) return 0.0;
  double beta = kf_beta(s);
  double umax = pow(1.0 - mV2 / s, beta / 2.0);
  double betaf = sqrt(1.0 - 4.0 * me * me / mV2);
  double prefac = (4.0 * kPi * kPi) * kAlpha * betaf * (3.0 / 2.0 - betaf * betaf / 2.0) / s * umax;
  double u = u0 * umax;
  double x1 = 1.0 - pow(u, 2.0 / beta);
  double x2 = mV2 / (x1 * s);
  if (!((x2 < 1.0) && (x1 > 0.0) && (u0 < 1.0))) return 0.0;
  double y = mV2 / s;
  double lumi = fl_kf(y / x1, beta) * fl_kf_scaled(x1, beta) * (1.0 / x1) * (2.0 / beta);
  return 2.0 * prefac * lumi;
}

// ---- the radiative-return integrand as the sampler evaluates it.  Everything that depends on the positron energy alone - s, beta
// (a logarithm), u_max (a pow), the prefactor (a square root and three divisions) - is computed ONCE per sample (SampleConst:
// b = beta, c = u_max, d = 2 x prefac, e = mV^2 / s) with the very expressions of ds_darkann; a trial is left with three powers.
// DarkAnn needs ~300 trials per accepted sample (the flux function is singular at the resonance), 1e9 trials per 2e4 config-3
// showers.  PB_DARKANN_FASTPOW: a^b as exp(b log a) through the constant-memory log (relative error ~ |b ln a| ulp <= 1e-14
// here, where libdevice's pow carries a double-double logarithm for <= 2 ulp); differences to ds_darkann are far inside the
// integrand's own conditioning (tests/test_gpu_probes.py::test_dark_dsigma_vs_reference_golden).
#ifndef PB_DARKANN_FASTPOW
#define PB_DARKANN_FASTPOW 1
#endif
__device__ __forceinline__ double ann_pow(double a, double b) {
#if PB_DARKANN_FASTPOW
  if (!(a > 0.0)) return (a == 0.0) ? (b > 0.0 ? 0.0 : 1.0 / 0.0) : a + b;      // pow(0, b); NaN in -> NaN out
  return exp(b * hot_log(a));
#else
  return pow(a, b);
#endif
}
__device__ __forceinline__ SampleConst darkann_const(const Material& M, double Ee) {
  const double me = kMe;
  SampleConst c{0.0, 0.0, 0.0, 0.0, 0.0};
  double mV2 = M.mV * M.mV;
  double s = 2.0 * me * (Ee + me);
  c.a = s;
  double beta = kf_beta(s);
  double umax = pow(1.0 - mV2 / s, beta / 2.0);
  double betaf = sqrt(1.0 - 4.0 * me * me / mV2);
  double prefac = (4.0 * kPi * kPi) * kAlpha * betaf * (3.0 / 2.0 - betaf * betaf / 2.0) / s * umax;
  c.b = beta; c.c = umax; c.d = 2.0 * prefac; c.e = mV2 / s;
  return c;
}
__device__ __forceinline__ double ds_darkann_c(const Material& M, const SampleConst& c, double Ee, double u0) {
  const double me = kMe;
  const double mV2 = M.mV * M.mV;
  const double s = 2.0 * me * (Ee + me);          // one multiply-add: cheaper than a fourth staged constant
  if (s < mV2) return 0.0;
  const double beta = c.b, umax = c.c;
  double u = u0 * umax;
  double x1 = 1.0 - ann_pow(u, 2.0 / beta);
  double x2 = mV2 / (x1 * s);
  if (!((x2 < 1.0) && (x1 > 0.0) && (u0 < 1.0))) return 0.0;
  double y = c.e;
  // fl_kf(y / x1, beta) * fl_kf_scaled(x1, beta) / x1 * (2 / beta), the two flux factors as in radiative_return.py:26-79
  double xa = y / x1;
  if (xa >= 1.0) xa = 1.0 - 1e-10;
  double fa = (beta / 16.0) * ((8.0 + 3.0 * beta) * ann_pow(1.0 - xa, beta / 2.0 - 1.0) - 4.0 * (1.0 + xa));
  double fb = (beta / 16.0) * ((8.0 + 3.0 * beta) - 4.0 * (1.0 + x1) * ann_pow(1.0 - x1, 1.0 - beta / 2.0));
  double lumi = fa * fb * (1.0 / x1) * (2.0 / beta);
  return c.d * lumi;
}

__device__ __forceinline__ double dsigma(const Material& M, int proc, double E, const double* x) {
  switch (proc) {
    case P_BREM: return ds_brem(M, E, kMe, x);
    case P_MUONBREM: return ds_brem(M, E, kMmu, x);
    case P_PAIRPROD: return ds_pairprod(M, E, x);
    case P_COMP: return ds_compton(M, E, 0.0, x[0]);
    case P_ANN: return ds_annihilation(E, 0.0, M.Eg_min, x[0]);
    case P_MOLLER: return ds_moller(E, M.Ee_min, x[0]);
    case P_BHABHA: return ds_bhabha(E, M.Ee_min, x[0]);
    case P_MUONE: return ds_muone(E, M.Ee_min, x[0]);
    case P_DARKBREM: return ds_darkbrem(M, E, kMe, x);
    case P_DARKMUONBREM: return ds_darkbrem(M, E, kMmu, x);
    case P_DARKANN: return ds_darkann(M, E, x[0]);
    case P_DARKCOMP: return ds_compton(M, E, M.mV, x[0]);
  }
  return 0.0;
}

// ---------------------------------------------------------------- energy loss, rotations, multiple scattering
struct V4 { double E, x, y, z; };

// particle.py:143-153
__device__ __forceinline__ V4 lose_energy(V4 p, double mass, double value) {
  double p30 = norm3_nofma(p.x, p.y, p.z);
  double Eu = p.E - value;
  if (Eu <= mass) Eu = mass;
  // no fma contraction: at Eu == mass the reference gets exactly 0 here and stops the particle (particle.py:149-153)
  double p3f = sqrt(__dsub_rn(__dmul_rn(Eu, Eu), __dmul_rn(mass, mass)));
  if (p3f > 0.0) { double r = p3f / p30; return V4{Eu, p.x * r, p.y * r, p.z * r}; }
  return V4{mass, 0.0, 0.0, 0.0};
}

// moliere.py:196-219, 265-281: Lynch-Dahl width of the Gaussian core, F = 0.98, z = 1.
// The reference forms the momentum in MeV as m_lepton * beta / sqrt(1 - beta^2) with beta = |p|/E; that is
// m_lepton * |p| / sqrt(E^2 - |p|^2) = m_lepton * |p| / mass, evaluated here without the gamma^2-amplified cancellation.
__device__ __forceinline__ double mcs_theta0(const Material& M, double t, double beta, double p_MeV) {
  const double F = 0.98;
  double pb = 1.0 / (p_MeV * beta);
  double chic2 = 0.157 * M.Z * (M.Z + 1) * (t / M.A) * (pb * pb);
  double za = M.Z * kAlpha / beta;
  double chia2 = 2.007e-5 * M.Z23 * (1.0 + 3.34 * (za * za)) / (p_MeV * p_MeV);
  double omega = chic2 / chia2;
  double v = 0.5 * omega / (1.0 - F);
  return sqrt(chic2 * ((1.0 + v) * log(1.0 + v) / v - 1) / (1.0 + F * F));
}

// moliere.py:350-400 (get_scattered_momentum_fast) with the rotation of moliere.py:287-348 in closed form.
// get_rotation_matrix builds Rb(b) Ra(a) from a = -/+atan|vy/vx| ..., b = ...atan|vx'/vz| with quadrant fix-ups (and a
// duplicated branch, SURVEY Q-13); every branch reduces to cos a = vx/r, sin a = -vy/r (r = hypot(vx, vy), so vx' = r)
// and cos b = vz/|v|, sin b = -vx'/|v|.  The axis-aligned special cases keep the reference's literal values.
// sign in {-1,+1}; radial = sqrt(z1^2 + z2^2) for the two unit normals; u_phi in [0,1).
__device__ __forceinline__ V4 mcs_apply(const Material& M, V4 p4, double pn, double t, double m_lepton, double mass,
                                        double sign, double radial, double u_phi) {
  double vx = p4.x, vy = p4.y, vz = p4.z;
  double beta = pn / p4.E;
  double ca, sa;
  if (vx != 0.0 && vy != 0.0) { double r = 1.0 / sqrt(vx * vx + vy * vy); ca = vx * r; sa = -vy * r; }
  else if (vy != 0.0) { ca = 0.0; sa = 1.0; }
  else { ca = 1.0; sa = 0.0; }
  double vxp = vx * ca - vy * sa;
  double cb, sb;
  if (vz != 0.0 && vxp != 0.0) { double r = 1.0 / sqrt(vxp * vxp + vz * vz); cb = vz * r; sb = -vxp * r; }
  else if (vxp > 0.0) { cb = 0.0; sb = -1.0; }
  else if (vxp < 0.0) { cb = 0.0; sb = 1.0; }
  else { cb = 1.0; sb = 0.0; }
  double th0 = mcs_theta0(M, t, beta, (m_lepton * 1e3) * (pn / mass));
  double theta = sign * (radial * th0) * M.rescale_mcs;
  double cth, sth, cph, sph;
  sincos(theta, &sth, &cth);
  sincospi(2.0 * u_phi, &sph, &cph);
  double q0 = pn * (sph * sth), q1 = pn * (-cph * sth), q2 = pn * cth;
  // R = Rb Ra = [[cb ca, -cb sa, sb], [sa, ca, 0], [-sb ca, sb sa, cb]] ; lab = R^T q
  V4 o;
  o.E = p4.E;
  o.x = (cb * ca) * q0 + sa * q1 + (-sb * ca) * q2;
  o.y = (-cb * sa) * q0 + ca * q1 + (sb * sa) * q2;
  o.z = sb * q0 + cb * q2;
  return o;
}

// Same physics as mcs_apply for the sub-step loop, with the algebra folded (2 divisions instead of 8):
//   beta^2 = pn^2/E^2,  p_MeV = Kp pn (Kp = 1e3 m_lepton / mass),  1/(p_MeV beta) = E / (Kp pn^2)
//   chic2 = C4 t / (p_MeV beta)^2,  omega = chic2/chia2 = Cw t / (beta^2 + c3),  v = omega / (2 (1-F))
// inv_pn = 1/pn is supplied by the caller (it also advances the position with it).  Results differ from mcs_apply
// by re-association only (~1e-15 relative).
__device__ __forceinline__ V4 mcs_fast(const Material& M, V4 p4, double pn, double inv_pn, double t, double iKp,
                                       double sign, double radial, double u_phi) {
  const double F = 0.98;
  double E = p4.E;
  double e_ip = E * inv_pn * inv_pn;                    // E / pn^2
  double ipb = e_ip * iKp;                              // 1 / (p_MeV beta)
  double chic2 = M.mcs_C4 * t * (ipb * ipb);
  double E2 = E * E;
  double omega = M.mcs_Cw * t * E2 * fast_rcp(pn * pn + M.mcs_c3 * E2);
  double v = omega * (0.5 / (1.0 - F));
  double th0 = fast_sqrt0(chic2 * ((1.0 + v) * hot_log(1.0 + v) * fast_rcp(v) - 1) * (1.0 / (1.0 + F * F)));
  double theta = sign * (radial * th0) * M.rescale_mcs;
  double vx = p4.x, vy = p4.y, vz = p4.z;
  double ca, sa, vxp;
  if (vx != 0.0 && vy != 0.0) { double pt2 = vx * vx + vy * vy; double r = fast_rsqrt(pt2); ca = vx * r; sa = -vy * r; vxp = pt2 * r; }
  else if (vy != 0.0) { ca = 0.0; sa = 1.0; vxp = -vy; }
  else { ca = 1.0; sa = 0.0; vxp = vx; }
  double cb, sb;
  if (vz != 0.0 && vxp != 0.0) { cb = vz * inv_pn; sb = -vxp * inv_pn; }
  else if (vxp > 0.0) { cb = 0.0; sb = -1.0; }
  else if (vxp < 0.0) { cb = 0.0; sb = 1.0; }
  else { cb = 1.0; sb = 0.0; }
  double cth, sth, cph, sph;
  hot_sincos(theta, &sth, &cth);
  hot_sincos_2pi(u_phi, &sph, &cph);
  double q0 = pn * (sph * sth), q1 = pn * (-cph * sth), q2 = pn * cth;
  V4 o;
  o.E = E;
  o.x = (cb * ca) * q0 + sa * q1 + (-sb * ca) * q2;
  o.y = (-cb * sa) * q0 + ca * q1 + (sb * sa) * q2;
  o.z = sb * q0 + cb * q2;
  return o;
}

// The MCS block's four draws: sign, two normals, azimuth.  CPython's gauss pair is
// (cos, sin)(2 pi u_a) * sqrt(-2 ln(1 - u_r)), so sqrt(z1^2 + z2^2) = sqrt(-2 ln(1 - u_r)) and u_a drops out: one Philox
// call gives (u_phi, u_r) and the sign comes from a spare bit (oracle/draws.py CounterDraws.mcs).
struct McsDraw { double sign, radial, uphi; };
__device__ __forceinline__ McsDraw mcs_draw(uint2 key, uint32_t index, uint32_t pc) {
  uint32_t sp;
  D2 a = draw2s(key, index, ST_MCS, 0, pc, sp);
  McsDraw d;
  d.sign = (sp & 1u) ? 1.0 : -1.0;
  d.uphi = a.a;
  d.radial = fast_sqrt0(-2.0 * hot_log(1.0 - a.b));
  return d;
}
__device__ __forceinline__ V4 mcs_scatter(const Material& M, V4 p4, double pn, double t, double m_lepton, double mass,
                                          const McsDraw& d) {
  return mcs_apply(M, p4, pn, t, m_lepton, mass, d.sign, d.radial, d.uphi);
}

// particle.py:176-185: rows of Rz(phi) Ry(theta) taking z-hat onto pf; applied as lab = R v
struct Rot { double r00, r01, r02, r10, r11, r12, r20, r22; };
__device__ __forceinline__ Rot rotation_to(V4 pf) {
  // theta goes through acos as in the reference: for a collimated particle acos(1 - eps) carries an error of ~1e-16 / theta that the
  // daughters inherit, and fidelity means inheriting the same one.  The azimuth is well conditioned: cos / sin of atan2(py, px) ARE
  // px / pt, py / pt to an ulp, which saves the atan2 and a full-range sincos (12 % of k_emit's instructions).
  double th = acos(pf.z / norm3_nofma(pf.x, pf.y, pf.z));
  double ct, st, cp = 1.0, sp = 0.0;
  hot_sincos(th, &st, &ct);
  double pt2 = pf.x * pf.x + pf.y * pf.y;
  if (pt2 > 0.0) { double r = fast_rsqrt(pt2); cp = pf.x * r; sp = pf.y * r; }     // atan2(0, 0) = 0
  return Rot{ct * cp, -sp, st * cp, ct * sp, cp, st * sp, -st, ct};
}
__device__ __forceinline__ V4 rotate(const Rot& R, V4 v) {
  return V4{v.E, R.r00 * v.x + R.r01 * v.y + R.r02 * v.z, R.r10 * v.x + R.r11 * v.y + R.r12 * v.z, R.r20 * v.x + R.r22 * v.z};
}

// ---------------------------------------------------------------- kinematics (parent along +z)
__device__ __forceinline__ double sq(double v) { return v * v; }

// kinematics.py:10-41 (e_to_egamma_fourvecs): a = outgoing lepton, b = photon
__device__ __forceinline__ void kin_brem(double ep, double ml, const double* x, double u_az, V4* a, V4* b) {
  double w = EGAMMA_MIN_KIN + x[0] * (ep - ml - EGAMMA_MIN_KIN);
  // (libdevice cos on purpose: sqrt(1 - ct^2) below amplifies the last bit of ct by 1 / theta^2, and fidelity to the reference's
  // value means the rounding closest to glibc's; the azimuths are well conditioned and go through the exact-reduction kernel)
  double ct = cos((x[1] + x[2]) / 2);
  double ctp = cos((x[1] - x[2]) * ep / (2 * (ep - w)));
  double epp = ep - w;
  double pp = sqrt(epp * epp - ml * ml);
  double sal, cal, sp, cp;
  hot_sincos_2pi(u_az, &sal, &cal);
  hot_sincos_2pi(x[3] - 0.5, &sp, &cp);            // ph = (x4 - 1/2) 2 pi
  double st = sqrt(1.0 - ct * ct), stp = sqrt(1.0 - ctp * ctp);
  *b = V4{w, w * cal * st, w * sal * st, w * ct};
  *a = V4{epp, pp * (sal * sp * stp + cal * (ctp * st - cp * ct * stp)), pp * (ctp * sal * st - (cp * ct * sal + cal * sp) * stp),
          pp * (ct * ctp + cp * st * stp)};
}

// kinematics.py:70-102 (gamma_to_epem_fourvecs): a = positron, b = electron
__device__ __forceinline__ void kin_pairprod(double w, const double* x, double u_az, V4* a, V4* b) {
  const double me = kMe;
  double epp = me + x[0] * (w - 2 * me);
  double ctp = cos(w * (x[1] + x[2]) / (2 * epp));
  double ctm = cos(w * (x[1] - x[2]) / (2 * (w - epp)));
  double epm = w - epp;
  double pm = sqrt(epm * epm - me * me), pp = sqrt(epp * epp - me * me);
  double sal, cal, spal, cpal;
  hot_sincos_2pi(u_az, &sal, &cal);                // al = 2 pi u
  hot_sincos_2pi(x[3] + u_az, &spal, &cpal);       // ph + al = 2 pi (x4 + u)
  double stp = sqrt(1.0 - ctp * ctp), stm = sqrt(1.0 - ctm * ctm);
  *a = V4{epp, pp * stp * cal, pp * stp * sal, pp * ctp};
  *b = V4{epm, pm * stm * cpal, pm * stm * spal, pm * ctm};
}

// kinematics.py:104-132 (compton_fourvecs): a = electron, b = photon / V
__device__ __forceinline__ void kin_compton(double Eg, double mV, double ct, double u_az, V4* a, V4* b) {
  const double me = kMe;
  double s = me * me + 2 * Eg * me;
  double rs = sqrt(s);
  double Ee0 = (s + me * me) / (2.0 * rs);
  double Ee = (s - mV * mV + me * me) / (2 * rs);
  double EV = (s + mV * mV - me * me) / (2 * rs);
  double pF = sqrt(Ee * Ee - me * me);
  double g0 = Ee0 / me;
  double b0 = 1.0 / g0 * sqrt(g0 * g0 - 1.0);
  double sp, cp;
  hot_sincos_2pi(u_az, &sp, &cp);
  double st = sqrt(1 - ct * ct);
  *a = V4{g0 * Ee + b0 * g0 * pF * ct, -pF * st * sp, -pF * st * cp, b0 * g0 * Ee + g0 * pF * ct};
  *b = V4{g0 * EV - b0 * g0 * pF * ct, pF * st * sp, pF * st * cp, b0 * g0 * EV - g0 * pF * ct};
}

// kinematics.py:301-334 (annihilation_fourvecs): a = photon, b = photon / V
__device__ __forceinline__ void kin_annihilation(double Ee, double mV, double ct, double u_az, V4* a, V4* b) {
  const double me = kMe;
  double s = 2 * me * (Ee + me);
  double rs = sqrt(s);
  double EeCM = rs / 2.0;
  double Eg = (s - mV * mV) / (2 * rs);
  double EV = (s + mV * mV) / (2 * rs);
  double pF = Eg;
  double g0 = EeCM / me;
  double b0 = 1.0 / g0 * sqrt(g0 * g0 - 1.0);
  double sp, cp;
  hot_sincos_2pi(u_az, &sp, &cp);
  double st = sqrt(1 - ct * ct);
  *a = V4{g0 * Eg - b0 * g0 * pF * ct, -pF * st * sp, -pF * st * cp, b0 * g0 * Eg - g0 * pF * ct};
  *b = V4{g0 * EV + b0 * g0 * pF * ct, pF * st * sp, pF * st * cp, b0 * g0 * EV + g0 * pF * ct};
}

// kinematics.py:213-237 (ee_to_ee_fourvecs): a = scattered e+-, b = struck electron
__device__ __forceinline__ void kin_ee(double Einc, double ct, double u_az, V4* a, V4* b) {
  const double me = kMe;
  double s = 2 * me * me + 2 * Einc * me;
  double Ee0 = sqrt(s) / 2.0;
  double pF = sqrt(Ee0 * Ee0 - me * me);
  double g0 = Ee0 / me;
  double b0 = 1.0 / g0 * sqrt(g0 * g0 - 1.0);
  double sp, cp;
  hot_sincos_2pi(u_az, &sp, &cp);
  double st = sqrt(1 - ct * ct);
  *a = V4{g0 * Ee0 + b0 * g0 * pF * ct, -pF * st * sp, -pF * st * cp, b0 * g0 * Ee0 + g0 * pF * ct};
  *b = V4{g0 * Ee0 - b0 * g0 * pF * ct, pF * st * sp, pF * st * cp, b0 * g0 * Ee0 - g0 * pF * ct};
}

// kinematics.py:239-265 (mue_to_mue_fourvecs): a = muon, b = electron
__device__ __forceinline__ void kin_mue(double Einc, double ct, double u_az, V4* a, V4* b) {
  const double me = kMe, mm = kMmu;
  double s = me * me + mm * mm + 2 * Einc * me;
  double rs = sqrt(s);
  double Ee0 = (s + me * me - mm * mm) / (2.0 * rs);
  double Em0 = (s + mm * mm - me * me) / (2.0 * rs);
  double pe = sqrt(Ee0 * Ee0 - me * me);
  double pm = sqrt(Em0 * Em0 - mm * mm);
  double g0 = Ee0 / me;
  double b0 = 1.0 / g0 * sqrt(g0 * g0 - 1.0);
  double sp, cp;
  hot_sincos_2pi(u_az, &sp, &cp);
  double st = sqrt(1 - ct * ct);
  *a = V4{g0 * Em0 + b0 * g0 * pm * ct, pm * st * sp, pm * st * cp, b0 * g0 * Em0 + g0 * pm * ct};
  *b = V4{g0 * Ee0 - b0 * g0 * pe * ct, -pe * st * sp, -pe * st * cp, b0 * g0 * Ee0 - g0 * pe * ct};
}

// kinematics.py:43-68 (l_to_lV_fourvecs): only the dark vector is used downstream
__device__ __forceinline__ V4 kin_darkbrem_V(double ep, double mV, const double* x, double u_az) {
  double w = x[0] * ep;
  double ct = 1 - pow(10.0, x[1]);
  double k = sqrt(w * w - mV * mV);
  double sal, cal;
  hot_sincos_2pi(u_az, &sal, &cal);
  double st = sqrt(1.0 - ct * ct);
  return V4{w, k * cal * st, k * sal * st, k * ct};
}

// kinematics.py:267-299 (radiative_return_fourvecs) with radiative_return.py:18-24 (boost): collinear ISR, no azimuth
__device__ __forceinline__ V4 kin_darkann_V(double Ee, double mV, double u0) {
  const double me = kMe;
  double s = 2.0 * me * (me + Ee);
  double beta = kf_beta(s);
  double umax = pow(1.0 - mV * mV / s, beta / 2.0);
  double x1 = 1.0 - pow(u0 * umax, 2.0 / beta);
  double x2 = mV * mV / (x1 * s);
  double rs = sqrt(s);
  double E1 = x1 * rs / 2.0, E2 = x2 * rs / 2.0;
  double v0 = E1 + E2, v3 = E1 - E2;                       // pV = p1 + p2 in the CM frame
  double p0 = rs / 2.0, p3 = -sqrt(s / 4.0 - me * me);     // CM four-momentum of the target electron
  double rsq = sqrt(p0 * p0 - p3 * p3);
  double b0 = (p0 * v0 - p3 * v3) / rsq;
  double c1 = (v0 + b0) / (rsq + p0);
  return V4{b0, 0.0, 0.0, v3 - c1 * p3};
}

// dark_shower.py:706-708 (electron_wave_function): hydrogenic |psi(p)|^2 p^2
__device__ __forceinline__ double electron_wave_function(double Zeff, double pe) {
  double lam = kAlpha * Zeff * kMe;
  double lam2 = lam * lam, d = pe * pe + lam2, d2 = d * d;
  return 32 / kPi * (lam2 * lam2 * lam) * pe * pe / (d2 * d2);
}

// kinematics.py:134-183 (compton_fourvecs_boundelectron): only the dark vector is returned
__device__ __forceinline__ V4 kin_compton_bound_V(double Eg, double mV, double ct, double Pe, double cte, double u1, double u2) {
  const double me = kMe;
  double s = me * me + 2 * Eg * (sqrt(me * me + Pe * Pe) - cte * Pe);
  double rs = sqrt(s);
  double Ee = (s - mV * mV + me * me) / (2 * rs);
  double EV = (s + mV * mV - me * me) / (2 * rs);
  double pF = sqrt(Ee * Ee - me * me);
  double bnum = sqrt(Eg * Eg + 2 * cte * Eg * Pe + Pe * Pe);
  double bden = Eg + sqrt(me * me + Pe * Pe);
  double b0 = bnum / bden;
  double g0 = 1.0 / sqrt(1.0 - b0 * b0);
  double sp, cp, se, ce;
  sincos(u1 * kTwoPi, &sp, &cp);
  double st = sqrt(1 - ct * ct);
  double EVLab = g0 * EV - b0 * g0 * pF * ct;
  double vx = pF * st * sp, vy = pF * st * cp, vz = b0 * g0 * EV - g0 * pF * ct;
  double ctz = (Eg + cte * Pe) / bnum;
  double stz = sqrt(1.0 - ctz * ctz);
  sincos(u2 * kTwoPi, &se, &ce);
  return V4{EVLab, ctz * ce * vx - se * vy + stz * ce * vz, ctz * se * vx + ce * vy + stz * se * vz, -stz * vx + ctz * vz};
}

// particle.py:187-256 (boost_matrix, two_body_decay, isotropic)
__device__ __forceinline__ void two_body_decay(V4 pf, double mX, double m1, double m2, double u_cos, double u_phi, V4* a, V4* b) {
  double E1 = (mX * mX - m2 * m2 + m1 * m1) / (2 * mX);
  double E2 = (mX * mX - m1 * m1 + m2 * m2) / (2 * mX);
  double pF = sqrt(E1 * E1 - m1 * m1);
  double c = -1.0 + 2.0 * u_cos;
  double sphi, cphi;
  sincos(kTwoPi * u_phi, &sphi, &cphi);
  double sth = sqrt(1 - c * c);
  double v1[4] = {E1, -pF * sth * sphi, -pF * sth * cphi, -pF * c};
  double v2[4] = {E2, pF * sth * sphi, pF * sth * cphi, pF * c};
  double gamma = pf.E / mX;
  double beta = (gamma == 1.0) ? 1.0 : sqrt(1.0 - 1.0 / (gamma * gamma));
  double pmag = norm3_nofma(pf.x, pf.y, pf.z);
  if (pmag == 0.0) {
    *a = V4{v1[0], v1[1], v1[2], v1[3]};
    *b = V4{v2[0], v2[1], v2[2], v2[3]};
    return;
  }
  double bv[3] = {beta * pf.x / pmag, beta * pf.y / pmag, beta * pf.z / pmag};
  double g1 = gamma - 1, b2 = beta * beta;
  double B[4][4];
  B[0][0] = gamma;
  for (int i = 0; i < 3; ++i) {
    B[0][i + 1] = B[i + 1][0] = gamma * bv[i];
    for (int j = 0; j < 3; ++j) B[i + 1][j + 1] = (i == j ? 1.0 : 0.0) + g1 * bv[i] * bv[j] / b2;
  }
  double o1[4], o2[4];
  for (int i = 0; i < 4; ++i) {
    o1[i] = B[i][0] * v1[0] + B[i][1] * v1[1] + B[i][2] * v1[2] + B[i][3] * v1[3];
    o2[i] = B[i][0] * v2[0] + B[i][1] * v2[1] + B[i][2] * v2[2] + B[i][3] * v2[3];
  }
  *a = V4{o1[0], o1[1], o1[2], o1[3]};
  *b = V4{o2[0], o2[1], o2[2], o2[3]};
}

}  // namespace pb
