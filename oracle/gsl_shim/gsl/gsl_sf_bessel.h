// Test infrastructure only (oracle/): minimal stand-in for <gsl/gsl_sf_bessel.h>
// so that the UNMODIFIED reference sources under /root/reference/src compile in
// an image without GSL.  Call sites: FSSW.cpp:1626-1633,1655,1675,1695.
// Implemented with the C++17 special functions of libstdc++ (std::cyl_bessel_k).
#ifndef ISS_ORACLE_GSL_SF_BESSEL_H
#define ISS_ORACLE_GSL_SF_BESSEL_H
#include <cmath>
static inline double gsl_sf_bessel_K1(double x) { return std::cyl_bessel_k(1.0, x); }
static inline double gsl_sf_bessel_Kn(int n, double x) {
    return std::cyl_bessel_k(static_cast<double>(n), x);
}
#endif
