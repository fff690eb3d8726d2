// QUADPACK QAGS (dqagse + dqk21 + dqelg + dqpsrt), restated for one CUDA thread per integral.
//
// Why it exists: DarkShower.__init__ builds its emission-weight and dRate/dE tables with scipy.integrate.quad
// (src/PETITE/dark_shower.py:337-399, 454-493; the cumulative interaction integrals of shower.py:298-320), i.e. QUADPACK's QAGS with
// epsabs = epsrel = 1.49e-8 and limit = 50 - and on these integrands (piecewise-linear tables with a hundred kinks, an exponential
// survival factor) QAGS regularly stops at its subdivision limit with errors of 1e-5 .. 1e-3.  The reference's tables therefore ARE the
// output of this particular algorithm, and reproducing them (tests: 1e-7 against the tables dumped from the reference constructor)
// means taking the same subdivision and extrapolation decisions, not integrating "better".  The routines follow the published
// Fortran (Piessens, de Doncker-Kapenga, Ueberhuber, Kahaner, QUADPACK 1983, public domain; the version SciPy wraps) statement by
// statement: 21-point Gauss-Kronrod rule, bisection of the interval with the largest error estimate, Wynn's epsilon algorithm.
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define PBQ_HD __host__ __device__
#else
#define PBQ_HD
#endif

namespace pbq {

constexpr int LIMIT = 50;                  // scipy.integrate.quad default
constexpr double EPMACH = 2.220446049250313e-16, UFLOW = 2.2250738585072014e-308, OFLOW = 1.7976931348623157e308;

// Gauss-Kronrod (10, 21) nodes and weights (dqk21.f)
#ifdef __CUDA_ARCH__
#define PBQ_CONST __constant__
#else
#define PBQ_CONST static const
#endif
PBQ_CONST double kWg[5] = {0.066671344308688137593568809893332, 0.149451349150580593145776339657697, 0.219086362515982043995534934228163,
                           0.269266719309996355091226921569469, 0.295524224714752870173815619188769};
PBQ_CONST double kXgk[11] = {0.995657163025808080735527280689003, 0.973906528517171720077964012084452, 0.930157491355708226001207180059508,
                             0.865063366688984510732096688423493, 0.780817726586416897063717578345042, 0.679409568299024406234327365114874,
                             0.562757134668604683339000099272694, 0.433395394129247190799265943165784, 0.294392862701460198131126603103866,
                             0.148874338981631210884826001129720, 0.0};
PBQ_CONST double kWgk[11] = {0.011694638867371874278064396062192, 0.032558162307964727478818972459390, 0.054755896574351996031381300244580,
                             0.075039674810919952767043140916190, 0.093125454583697605535065465083366, 0.109387158802297641899210590325805,
                             0.123491976262065851077958109585166, 0.134709217311473325928054001771707, 0.142775938577060080797094273138717,
                             0.147739104901338491374841515972068, 0.149445554002916905664936468389821};

template <class F>
PBQ_HD inline void qk21(const F& f, double a, double b, double& result, double& abserr, double& resabs, double& resasc) {
  double fv1[10], fv2[10];
  const double centr = 0.5 * (a + b), hlgth = 0.5 * (b - a), dhlgth = fabs(hlgth);
  double resg = 0.0;
  const double fc = f(centr);
  double resk = kWgk[10] * fc;
  resabs = fabs(resk);
  for (int j = 0; j < 5; ++j) {
    const int jtw = 2 * j + 1;
    const double absc = hlgth * kXgk[jtw];
    const double f1 = f(centr - absc), f2 = f(centr + absc);
    fv1[jtw] = f1; fv2[jtw] = f2;
    const double fsum = f1 + f2;
    resg += kWg[j] * fsum;
    resk += kWgk[jtw] * fsum;
    resabs += kWgk[jtw] * (fabs(f1) + fabs(f2));
  }
  for (int j = 0; j < 5; ++j) {
    const int jtwm1 = 2 * j;
    const double absc = hlgth * kXgk[jtwm1];
    const double f1 = f(centr - absc), f2 = f(centr + absc);
    fv1[jtwm1] = f1; fv2[jtwm1] = f2;
    const double fsum = f1 + f2;
    resk += kWgk[jtwm1] * fsum;
    resabs += kWgk[jtwm1] * (fabs(f1) + fabs(f2));
  }
  const double reskh = resk * 0.5;
  resasc = kWgk[10] * fabs(fc - reskh);
  for (int j = 0; j < 10; ++j) resasc += kWgk[j] * (fabs(fv1[j] - reskh) + fabs(fv2[j] - reskh));
  result = resk * hlgth;
  resabs *= dhlgth;
  resasc *= dhlgth;
  abserr = fabs((resk - resg) * hlgth);
  if (resasc != 0.0 && abserr != 0.0) abserr = resasc * fmin(1.0, pow(200.0 * abserr / resasc, 1.5));
  if (resabs > UFLOW / (50.0 * EPMACH)) abserr = fmax((EPMACH * 50.0) * resabs, abserr);
}

// dqpsrt: keeps iord (1-based indices, as in the Fortran) sorted by decreasing error estimate
PBQ_HD inline void qpsrt(int limit, int last, int& maxerr, double& ermax, const double* elist, int* iord, int& nrmax) {
  // arrays are used 1-based: element k lives at [k]
  if (last <= 2) { iord[1] = 1; iord[2] = 2; maxerr = iord[nrmax]; ermax = elist[maxerr]; return; }
  const double errmax = elist[maxerr];
  if (nrmax != 1) {
    const int ido = nrmax - 1;
    for (int i = 1; i <= ido; ++i) {
      const int isucc = iord[nrmax - 1];
      if (errmax <= elist[isucc]) break;
      iord[nrmax] = isucc;
      nrmax -= 1;
    }
  }
  int jupbn = last;
  if (last > (limit / 2 + 2)) jupbn = limit + 3 - last;
  const double errmin = elist[last];
  const int jbnd = jupbn - 1, ibeg = nrmax + 1;
  int i = ibeg;
  bool found = false;
  if (ibeg <= jbnd) {
    for (i = ibeg; i <= jbnd; ++i) {
      const int isucc = iord[i];
      if (errmax >= elist[isucc]) { found = true; break; }
      iord[i - 1] = isucc;
    }
  }
  if (!found) {
    iord[jbnd] = maxerr;
    iord[jupbn] = last;
  } else {
    iord[i - 1] = maxerr;
    int k = jbnd;
    bool placed = false;
    for (int j = i; j <= jbnd; ++j) {
      const int isucc = iord[k];
      if (errmin < elist[isucc]) { iord[k + 1] = last; placed = true; break; }
      iord[k + 1] = isucc;
      k -= 1;
    }
    if (!placed) iord[i] = last;
  }
  maxerr = iord[nrmax];
  ermax = elist[maxerr];
}

// dqelg: epsilon algorithm (epstab 1-based, 52 + 2 elements; res3la 1-based, 3 elements)
PBQ_HD inline void qelg(int& n, double* epstab, double& result, double& abserr, double* res3la, int& nres) {
  nres += 1;
  abserr = OFLOW;
  result = epstab[n];
  if (n >= 3) {
    const int limexp = 50;
    epstab[n + 2] = epstab[n];
    const int newelm = (n - 1) / 2;
    epstab[n] = OFLOW;
    const int num = n;
    int k1 = n;
    bool converged = false;
    for (int i = 1; i <= newelm; ++i) {
      const int k2 = k1 - 1, k3 = k1 - 2;
      double res = epstab[k1 + 2];
      const double e0 = epstab[k3], e1 = epstab[k2], e2 = res;
      const double e1abs = fabs(e1);
      const double delta2 = e2 - e1, err2 = fabs(delta2), tol2 = fmax(fabs(e2), e1abs) * EPMACH;
      const double delta3 = e1 - e0, err3 = fabs(delta3), tol3 = fmax(e1abs, fabs(e0)) * EPMACH;
      if (!(err2 > tol2 || err3 > tol3)) {          // e0, e1, e2 equal to machine accuracy: convergence
        result = res;
        abserr = err2 + err3;
        converged = true;
        break;
      }
      const double e3 = epstab[k1];
      epstab[k1] = e1;
      const double delta1 = e1 - e3, err1 = fabs(delta1), tol1 = fmax(e1abs, fabs(e3)) * EPMACH;
      if (err1 <= tol1 || err2 <= tol2 || err3 <= tol3) { n = i + i - 1; break; }
      const double ss = 1.0 / delta1 + 1.0 / delta2 - 1.0 / delta3;
      const double epsinf = fabs(ss * e1);
      if (!(epsinf > 1.0e-4)) { n = i + i - 1; break; }
      res = e1 + 1.0 / ss;
      epstab[k1] = res;
      k1 -= 2;
      const double error = err2 + fabs(res - e2) + err3;
      if (error > abserr) continue;
      abserr = error;
      result = res;
    }
    if (!converged) {
      if (n == limexp) n = 2 * (limexp / 2) - 1;
      int ib = ((num / 2) * 2 == num) ? 2 : 1;
      const int ie = newelm + 1;
      for (int i = 1; i <= ie; ++i) { const int ib2 = ib + 2; epstab[ib] = epstab[ib2]; ib = ib2; }
      if (num != n) {
        int indx = num - n + 1;
        for (int i = 1; i <= n; ++i) { epstab[i] = epstab[indx]; indx += 1; }
      }
      if (nres < 4) {
        res3la[nres] = result;
        abserr = OFLOW;
        abserr = fmax(abserr, 5.0 * EPMACH * fabs(result));
        return;
      }
    }
    abserr = fabs(result - res3la[3]) + fabs(result - res3la[2]) + fabs(result - res3la[1]);
    res3la[1] = res3la[2];
    res3la[2] = res3la[3];
    res3la[3] = result;
  }
  abserr = fmax(abserr, 5.0 * EPMACH * fabs(result));
}

struct QagsOut { double result, abserr; int neval, ier, last; };

// dqagse with limit = LIMIT
template <class F>
PBQ_HD inline QagsOut qags(const F& f, double a, double b, double epsabs, double epsrel) {
  const int limit = LIMIT;
  double alist[LIMIT + 1], blist[LIMIT + 1], rlist[LIMIT + 1], elist[LIMIT + 1], rlist2[55], res3la[4] = {0.0, 0.0, 0.0, 0.0};
  int iord[LIMIT + 2];
  QagsOut o{0.0, 0.0, 0, 0, 0};
  int ier = 0, last = 0;
  double result = 0.0, abserr = 0.0;
  alist[1] = a; blist[1] = b; rlist[1] = 0.0; elist[1] = 0.0;
  if (epsabs <= 0.0 && epsrel < fmax(50.0 * EPMACH, 0.5e-28)) { o.ier = 6; return o; }
  int ierro = 0;
  double defabs, resabs;
  qk21(f, a, b, result, abserr, defabs, resabs);
  double dres = fabs(result);
  double errbnd = fmax(epsabs, epsrel * dres);
  last = 1;
  rlist[1] = result; elist[1] = abserr; iord[1] = 1;
  if (abserr <= 100.0 * EPMACH * defabs && abserr > errbnd) ier = 2;
  if (limit == 1) ier = 1;
  if (ier != 0 || (abserr <= errbnd && abserr != resabs) || abserr == 0.0) {
    o.result = result; o.abserr = abserr; o.ier = ier; o.last = last; o.neval = 42 * last - 21;
    return o;
  }
  rlist2[1] = result;
  double errmax = abserr;
  int maxerr = 1;
  double area = result, errsum = abserr;
  abserr = OFLOW;
  int nrmax = 1, nres = 0, numrl2 = 2, ktmin = 0;
  bool extrap = false, noext = false;
  int iroff1 = 0, iroff2 = 0, iroff3 = 0;
  int ksgn = -1;
  if (dres >= (1.0 - 50.0 * EPMACH) * defabs) ksgn = 1;
  double small = 0.0, erlarg = 0.0, ertest = 0.0, correc = 0.0, erlast = 0.0;
  int exit_to = 100;                         // 100: "set final result", 115: "compute global integral sum"
  for (last = 2; last <= limit; ++last) {
    const double a1 = alist[maxerr], b1 = 0.5 * (alist[maxerr] + blist[maxerr]), a2 = b1, b2 = blist[maxerr];
    erlast = errmax;
    double area1, error1, area2, error2, defab1, defab2, ra;
    qk21(f, a1, b1, area1, error1, ra, defab1);
    qk21(f, a2, b2, area2, error2, ra, defab2);
    const double area12 = area1 + area2, erro12 = error1 + error2;
    errsum = errsum + erro12 - errmax;
    area = area + area12 - rlist[maxerr];
    if (defab1 != error1 && defab2 != error2) {
      if (fabs(rlist[maxerr] - area12) <= 1.0e-5 * fabs(area12) && erro12 >= 0.99 * errmax) { if (extrap) iroff2 += 1; else iroff1 += 1; }
      if (last > 10 && erro12 > errmax) iroff3 += 1;
    }
    rlist[maxerr] = area1;
    rlist[last] = area2;
    errbnd = fmax(epsabs, epsrel * fabs(area));
    if (iroff1 + iroff2 >= 10 || iroff3 >= 20) ier = 2;
    if (iroff2 >= 5) ierro = 3;
    if (last == limit) ier = 1;
    if (fmax(fabs(a1), fabs(b2)) <= (1.0 + 100.0 * EPMACH) * (fabs(a2) + 1000.0 * UFLOW)) ier = 4;
    if (error2 > error1) {
      alist[maxerr] = a2; alist[last] = a1; blist[last] = b1;
      rlist[maxerr] = area2; rlist[last] = area1;
      elist[maxerr] = error2; elist[last] = error1;
    } else {
      alist[last] = a2; blist[maxerr] = b1; blist[last] = b2;
      elist[maxerr] = error1; elist[last] = error2;
    }
    qpsrt(limit, last, maxerr, errmax, elist, iord, nrmax);
    if (errsum <= errbnd) { exit_to = 115; break; }
    if (ier != 0) break;
    if (last == 2) { small = fabs(b - a) * 0.375; erlarg = errsum; ertest = errbnd; rlist2[2] = area; continue; }
    if (noext) continue;
    erlarg -= erlast;
    if (fabs(b1 - a1) > small) erlarg += erro12;
    if (!extrap) {
      if (fabs(blist[maxerr] - alist[maxerr]) > small) continue;
      extrap = true;
      nrmax = 2;
    }
    if (ierro != 3 && erlarg > ertest) {
      const int id = nrmax;
      int jupbnd = last;
      if (last > (2 + limit / 2)) jupbnd = limit + 3 - last;
      bool again = false;
      for (int k = id; k <= jupbnd; ++k) {
        maxerr = iord[nrmax];
        errmax = elist[maxerr];
        if (fabs(blist[maxerr] - alist[maxerr]) > small) { again = true; break; }
        nrmax += 1;
      }
      if (again) continue;
    }
    numrl2 += 1;
    rlist2[numrl2] = area;
    double reseps, abseps;
    qelg(numrl2, rlist2, reseps, abseps, res3la, nres);
    ktmin += 1;
    if (ktmin > 5 && abserr < 1.0e-3 * errsum) ier = 5;
    if (abseps < abserr) {
      ktmin = 0;
      abserr = abseps;
      result = reseps;
      correc = erlarg;
      ertest = fmax(epsabs, epsrel * fabs(reseps));
      if (abserr <= ertest) break;
    }
    if (numrl2 == 1) noext = true;
    if (ier == 5) break;
    maxerr = iord[1];
    errmax = elist[maxerr];
    nrmax = 1;
    extrap = false;
    small *= 0.5;
    erlarg = errsum;
  }
  if (last > limit) last = limit;             // (the loop always leaves through a break at last == limit at the latest)
  bool sum_up = exit_to == 115;
  if (!sum_up) {
    // label 100
    if (abserr == OFLOW) sum_up = true;
    else {
      bool test_div = true;
      if (ier + ierro != 0) {
        if (ierro == 3) abserr += correc;
        if (ier == 0) ier = 3;
        if (result != 0.0 && area != 0.0) {
          if (abserr / fabs(result) > errsum / fabs(area)) sum_up = true;
        } else {
          if (abserr > errsum) sum_up = true;
          else if (area == 0.0) test_div = false;
        }
      }
      if (!sum_up && test_div) {
        if (!(ksgn == -1 && fmax(fabs(result), fabs(area)) <= defabs * 0.01))
          if (0.01 > (result / area) || (result / area) > 100.0 || errsum > fabs(area)) ier = 6;
      }
    }
  }
  if (sum_up) {
    result = 0.0;
    for (int k = 1; k <= last; ++k) result += rlist[k];
    abserr = errsum;
  }
  if (ier > 2) ier -= 1;
  o.result = result; o.abserr = abserr; o.ier = ier; o.last = last; o.neval = 42 * last - 21;
  return o;
}

}  // namespace pbq
