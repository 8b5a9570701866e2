// Test infrastructure only (oracle/): stand-in for <gsl/gsl_sf_expint.h>.
// Call sites: FSSW.cpp:1637-1639,1715-1717.  E_n(x) for n>=2, x>0 by the
// modified-Lentz continued fraction (x>=1) or the power series (x<1).
#ifndef ISS_ORACLE_GSL_SF_EXPINT_H
#define ISS_ORACLE_GSL_SF_EXPINT_H
#include <cmath>
#include <limits>
static inline double iss_shim_expint_En(int n, double x) {
    const double eps = 1e-16, tiny = 1e-300, euler = 0.57721566490153286061;
    const int nm1 = n - 1;
    if (x == 0.0) return 1.0/nm1;
    if (x > 1.0) {
        double b = x + n, c = 1.0/tiny, d = 1.0/b, h = d;
        for (int i = 1; i <= 10000; i++) {
            double a = -1.0*i*(nm1 + i);
            b += 2.0;
            d = 1.0/(a*d + b);
            c = b + a/c;
            double del = c*d;
            h *= del;
            if (std::fabs(del - 1.0) < eps) break;
        }
        return h*std::exp(-x);
    }
    double ans = (nm1 != 0 ? 1.0/nm1 : -std::log(x) - euler);
    double fact = 1.0;
    for (int i = 1; i <= 10000; i++) {
        fact *= -x/i;
        double del;
        if (i != nm1) {
            del = -fact/(i - nm1);
        } else {
            double psi = -euler;
            for (int ii = 1; ii <= nm1; ii++) psi += 1.0/ii;
            del = fact*(-std::log(x) + psi);
        }
        ans += del;
        if (std::fabs(del) < std::fabs(ans)*eps) break;
    }
    return ans;
}
static inline double gsl_sf_expint_E2(double x) { return iss_shim_expint_En(2, x); }
static inline double gsl_sf_expint_En(int n, double x) { return iss_shim_expint_En(n, x); }
#endif
