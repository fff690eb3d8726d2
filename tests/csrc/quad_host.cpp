// Host build of petite_b200/csrc/quadpack.cuh with the SAME integrands as k_quad (engine.cu), for the CPU test that pins the QAGS
// restatement to scipy.integrate.quad (tests/test_quadpack_cpu.py).  Test infrastructure only.
#include <math.h>
#include <stdint.h>
#include "../../include/petite_b200.h"
#include "../../petite_b200/csrc/quadpack.cuh"

struct Tab { const double* x; const double* y; int n; double fill; };
static double lin(const Tab& T, double v) {
  int lo = 0, hi = T.n;
  while (lo < hi) { int mid = (lo + hi) >> 1; if (T.x[mid] < v) lo = mid + 1; else hi = mid; }
  hi = lo < 1 ? 1 : (lo > T.n - 1 ? T.n - 1 : lo);
  lo = hi - 1;
  volatile double slope = (T.y[hi] - T.y[lo]) / (T.x[hi] - T.x[lo]);
  volatile double prod = slope * (v - T.x[lo]);
  double out = prod + T.y[lo];
  return (v < T.x[0] || v > T.x[T.n - 1]) ? T.fill : out;
}
struct Integrand {
  const Tab* t; pb_quad_call c; double dEdx_m;
  double operator()(double E) const {
    if (c.kind == 0) return lin(t[c.tab], E);
    double v = pow(10.0, lin(t[c.tab], log10(E)));
    if (c.cut && v < 1.0e-18) return 0.0;
    double d = 0.0;
    for (int k = 0; k < 3; ++k) if (c.surv[k] >= 0) d = d + (lin(t[c.surv[k]], c.Ei) - lin(t[c.surv[k]], E));
    if (d < 0.0 || E > c.Ei) return 0.0;
    const double dEdx_cm = dEdx_m * 0.01;
    return v / dEdx_cm * exp(-d / dEdx_m / 0.01);
  }
};
extern "C" int quad_host_batch(int n_tabs, const int32_t* tab_n, const double* const* tab_x, const double* const* tab_y, const double* tab_fill,
                               double dEdx_m, const pb_quad_call* calls, int64_t n, double* result, double* abserr, int32_t* ier) {
  Tab tabs[64];
  if (n_tabs > 64) return -2;
  for (int k = 0; k < n_tabs; ++k) tabs[k] = Tab{tab_x[k], tab_y[k], tab_n[k], tab_fill[k]};
  for (int64_t i = 0; i < n; ++i) {
    Integrand f{tabs, calls[i], dEdx_m};
    pbq::QagsOut o = pbq::qags(f, f.c.a, f.c.b, 1.49e-8, 1.49e-8);
    result[i] = o.result;
    if (abserr) abserr[i] = o.abserr;
    if (ier) ier[i] = o.ier * 1000 + o.last;
  }
  return 0;
}
