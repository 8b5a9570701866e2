// Test infrastructure only (oracle/): stand-in for <gsl/gsl_randist.h>
// (FSSW.cpp:282,291,297).
#ifndef ISS_ORACLE_GSL_RANDIST_H
#define ISS_ORACLE_GSL_RANDIST_H
#include <random>
#include "gsl_rng.h"
static inline unsigned int gsl_ran_poisson(gsl_rng *r, double mu) {
    std::poisson_distribution<long> dist(mu);
    return static_cast<unsigned int>(dist(r->engine));
}
static inline unsigned int gsl_ran_negative_binomial(gsl_rng *r, double p, double n) {
    // GSL definition: X ~ Poisson(Y), Y ~ Gamma(n, (1-p)/p)
    std::gamma_distribution<double> gam(n, (1.0 - p)/p);
    double y = gam(r->engine);
    std::poisson_distribution<long> dist(y);
    return static_cast<unsigned int>(dist(r->engine));
}
#endif
