// Test infrastructure only (oracle/): stand-in for <gsl/gsl_sf_lambert.h>
// (legacy EmissionFunctionArray only: emissionfunction.cpp:3868,3946).
#ifndef ISS_ORACLE_GSL_SF_LAMBERT_H
#define ISS_ORACLE_GSL_SF_LAMBERT_H
#include <cmath>
static inline double gsl_sf_lambert_W0(double x) {
    if (x == 0.0) return 0.0;
    double w = (x < 1.0) ? x/(1.0 + x) : std::log(x) - std::log(std::log(x) + 1.0);
    for (int i = 0; i < 100; i++) {   // Halley iteration
        double ew = std::exp(w), f = w*ew - x;
        double dw = f/(ew*(w + 1.0) - (w + 2.0)*f/(2.0*w + 2.0));
        w -= dw;
        if (std::fabs(dw) < 1e-15*(1.0 + std::fabs(w))) break;
    }
    return w;
}
#endif
