// Test infrastructure only (oracle/): stand-in for <gsl/gsl_rng.h>
// (FSSW.cpp:169-172,227).  gsl_rng_default is mt19937 in GSL; here a
// std::mt19937 is wrapped, so the multiplicity stream is statistically but not
// bit-wise the one a real-GSL build would give.
#ifndef ISS_ORACLE_GSL_RNG_H
#define ISS_ORACLE_GSL_RNG_H
#include <random>
struct gsl_rng_type { const char *name; };
struct gsl_rng { std::mt19937 engine; };
static const gsl_rng_type iss_shim_mt19937_type = {"mt19937"};
static const gsl_rng_type *gsl_rng_default = &iss_shim_mt19937_type;
static inline const gsl_rng_type *gsl_rng_env_setup() { return gsl_rng_default; }
static inline gsl_rng *gsl_rng_alloc(const gsl_rng_type *) { return new gsl_rng; }
static inline void gsl_rng_set(gsl_rng *r, unsigned long seed) { r->engine.seed(seed); }
static inline void gsl_rng_free(gsl_rng *r) { delete r; }
#endif
