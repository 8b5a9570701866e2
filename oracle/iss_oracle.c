/*
 * iss_oracle.c -- TEST INFRASTRUCTURE ONLY.  CPU restatement, in plain C, of the sampling half of
 * the reference's hot path (chunshen1987/iSS), used by tests/, __graft_entry__.smoke() and the
 * cpu_baseline leg of bench.py as the CHECKER of the CUDA engine.  Nothing in the product links,
 * loads or calls this file.
 *
 * What is restated, and from where (paths relative to the reference tree):
 *   momentum tables + |p| sampler  src/BosonMomentumSampler.cpp:7-87, src/FermionMomentumSampler.cpp:7-93,
 *                                  src/MomentumSamplerShell.cpp:8-48, src/MomentumSamplerBase.cpp:20-93
 *   cell choice                    src/RandomVariable1DArray.cpp:25-67, src/arsenal.cpp:644-678
 *   accept/reject, delta f, boost  src/FSSW.cpp:1795-1966, 2001-2010
 *   emit                           src/FSSW.cpp:1969-1996 and the loop at :970-1050
 *   decays                         src/particle_decay.cpp:265-546, src/FSSW.cpp:1746-1779
 *
 * What is NOT the reference's and is restated from the engine's own design ("random
 * numbers"): the reference draws from one shared std::mt19937 in program order and from
 * gsl_ran_poisson (GSL is a third-party dependency absent from the reference tree, version
 * unpinned: CMakeLists.txt:16); the engine uses Philox4x32-10 (Salmon, Moraes, Dror, Shaw, SC'11)
 * keyed by (seed; stream, species, event, draw) and an exact Poisson inversion from the mode.
 * Those two pieces are restated here independently so that, GIVEN identical yields, the integer
 * bookkeeping is bit-exact and every hadron can be compared one to one.
 *
 * The last part of the file restates the legacy "conventional" sampler of EmissionFunctionArray
 * (MC_sampling = 2): estimate_maximum and the sampling loops; its yields are restated in numpy
 * (oracle/legacy_oracle.py).  Pinned by tests/test_legacy_cpu.py against dumps of the compiled
 * reference (oracle/ref_driver.cpp `legacy`: yields to 1e-12, maxima to 1e-12) and against the
 * histograms of the reference's own MC_sampling = 2 samples.
 *
 * Parity pinning: the reference algorithms in this file are pinned against the compiled reference
 * itself by tests/test_oracle_cpu.py (|p| spectra of MomentumSamplerShell and decay daughters of
 * particle_decay dumped by oracle/ref_driver.cpp into tests/golden/, and the reference's own
 * particle_samples.bin histograms).
 */
#define _GNU_SOURCE
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define HBARC 0.197327053 /* src/data_struct.h:9 */
#define NFIELD 28

enum { F_TAU = 0, F_X, F_Y, F_ETA, F_DA0, F_DA1, F_DA2, F_DA3, F_UT, F_UX, F_UY, F_UZ,
       F_E, F_T, F_P, F_NB, F_MUB, F_MUS, F_MUQ, F_BULK, F_PIXX, F_PIXY, F_PIXZ, F_PIYY, F_PIYZ,
       F_QX, F_QY, F_QZ };

typedef struct {
    int32_t pid, gspin, baryon, strange, charge, sign, decay_idx, reserved;
    double mass;
} o_species;

typedef struct {
    int32_t hydro_mode, include_shear, include_bulk, include_diff, bulk_kind, model, lcc, reserved;
    double para1, y_LB, y_RB;
} o_options;

typedef struct {
    int32_t pid;
    float mass, E, px, py, pz, t, x, y, z;
} o_hadron;

typedef struct {
    int32_t pid, stable, n_channels, first_channel, baryon, strange, charge, reserved;
    double mass, width;
} o_dspecies;

typedef struct {
    int32_t n_part;
    int32_t daughter[5];
    double br;
} o_dchannel;

/* ------------------------------------------------------------------ Philox4x32-10 */
static void philox(const uint32_t c[4], const uint32_t k[2], uint32_t out[4]) {
    uint32_t c0 = c[0], c1 = c[1], c2 = c[2], c3 = c[3], k0 = k[0], k1 = k[1];
    for (int r = 0; r < 10; r++) {
        const uint64_t p0 = (uint64_t)0xD2511F53u*c0, p1 = (uint64_t)0xCD9E8D57u*c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

typedef struct {
    uint32_t ctr[4], key[2], buf[4];
    int pos;
} o_stream;

enum { STREAM_MULT = 1, STREAM_SAMPLE = 2, STREAM_DECAY = 3 };

static void stream_init(o_stream *s, uint64_t seed, uint32_t stream, uint32_t species, uint32_t event,
                        uint32_t draw) {
    s->key[0] = (uint32_t)seed;
    s->key[1] = (uint32_t)(seed >> 32);
    s->ctr[0] = 0;
    s->ctr[1] = draw;
    s->ctr[2] = event;
    s->ctr[3] = (stream << 24) | (species & 0xFFFFFFu);
    s->pos = 4;
}

/* the four words of each Philox block are consumed in order */
static uint32_t stream_word(o_stream *s) {
    if (s->pos == 4) {
        philox(s->ctr, s->key, s->buf);
        s->ctr[0]++;
        s->pos = 0;
    }
    return s->buf[s->pos++];
}

/* 53-bit uniform in [0,1) from two words: 27 high bits of the first, 26 high bits of the second */
static double stream_next(o_stream *s) {
    const uint32_t a = stream_word(s);
    const uint32_t b = stream_word(s);
    const uint64_t v = ((uint64_t)(a >> 5) << 26) | (uint64_t)(b >> 6);
    return (double)v*(1.0/9007199254740992.0);
}

/* Block-granular use of the sampling streams (engine design, DESIGN.md "Random numbers"): every
 * decision point consumes one whole Philox block; unused words are discarded.
 *   cell choice       (w0,w1) -> 53-bit uniform
 *   |p| proposal      (w0,w1) -> r (53 bit), w2 -> inner accept (32 bit), w3 -> phi (32 bit)
 *   direction/accept  w0 -> cos(theta) (32 bit), w1 -> accept (32 bit)   [after the inner accept]
 *   rapidity          w0 -> y (32 bit)                                   [boost-invariant]      */
static void stream_block(o_stream *s, uint32_t w[4]) {
    philox(s->ctr, s->key, w);
    s->ctr[0]++;
}
static double w53(uint32_t a, uint32_t b) {
    const uint64_t v = ((uint64_t)(a >> 5) << 26) | (uint64_t)(b >> 6);
    return (double)v*(1.0/9007199254740992.0);
}
static double w32(uint32_t x) { return ((double)x + 0.5)*(1.0/4294967296.0); }

void oracle_philox(const uint32_t *ctr, const uint32_t *key, uint32_t *out) { philox(ctr, key, out); }

/* uniforms of one stream, for the RNG known-answer tests */
void oracle_stream_uniforms(uint64_t seed, uint32_t stream, uint32_t species, uint32_t event,
                            uint32_t draw, int n, double *out) {
    o_stream s;
    stream_init(&s, seed, stream, species, event, draw);
    for (int i = 0; i < n; i++) out[i] = stream_next(&s);
}

/* ------------------------------------------------------------------ multiplicities
 * Poisson by inversion outward from the mode m = floor(lambda) (engine design); pmode = pmf(m).
 * lambda < 1e-15 -> 0 as FSSW::determine_number_to_sample (src/FSSW.cpp:293-298).
 * Model 1: floor + Bernoulli(fraction) (src/FSSW.cpp:269-274). */
static int64_t poisson_mode(double lambda, double pmode, double u) {
    if (lambda < 1e-15) return 0;
    const int64_t m = (int64_t)lambda;
    u -= pmode;
    if (u < 0.0) return m;
    int64_t hi = m, lo = m;
    double phi = pmode, plo = pmode;
    for (;;) {
        hi++;
        phi = phi*(lambda/(double)hi);
        u -= phi;
        if (u < 0.0) return hi;
        if (lo > 0) {
            plo = plo*((double)lo/lambda);
            lo--;
            u -= plo;
            if (u < 0.0) return lo;
        }
        if (phi < 1e-300 && (lo == 0 || plo < 1e-300)) return m;
    }
}

/* Models 10 / 20 (src/FSSW.cpp:275-292): gsl_ran_negative_binomial(p = 1/(1+para1), k), i.e.
 * X ~ Poisson(Y) with Y ~ Gamma(shape k, scale para1) (GSL's published definition; GSL is not in
 * the reference tree).  Gamma by Marsaglia & Tsang (2000) + U^(1/k) boost, Box-Muller normals,
 * Poisson by the inversion above; uniforms taken from the stream in this order. */
static double gamma_draw(o_stream *r, double shape) {
    double boost = 1.0;
    if (shape < 1.0) {
        boost = pow(stream_next(r), 1.0/shape);
        shape += 1.0;
    }
    const double d = shape - 1.0/3.0, c = 1.0/sqrt(9.0*d);
    for (int it = 0; it < 1000; it++) {
        const double u1 = stream_next(r), u2 = stream_next(r);
        const double x = sqrt(-2.0*log(u1 > 0.0 ? u1 : 1e-300))*cos(6.283185307179586*u2);
        const double t = 1.0 + c*x;
        if (t <= 0.0) continue;
        const double v = t*t*t;
        const double u = stream_next(r);
        if (log(u > 0.0 ? u : 1e-300) < 0.5*x*x + d - d*v + d*log(v)) return boost*d*v;
    }
    return boost*d;
}

static int64_t nbd_draw(o_stream *r, double k, double scale) {
    const double y = scale*gamma_draw(r, k);
    if (y < 1e-15) return 0;
    const double m = floor(y);
    const double pm = exp(m*log(y) - y - lgamma(m + 1.0));
    return poisson_mode(y, pm, stream_next(r));
}

void oracle_multiplicities(const double *lambda, const double *pmode, const o_species *sp, int ns,
                           int64_t nev, int64_t ev_begin, uint64_t seed, int model, double para1,
                           int lcc, int64_t *mult /*[nev][ns]*/, int64_t *out_count /*[nev][ns]*/) {
    for (int64_t ev = 0; ev < nev; ev++)
        for (int s = 0; s < ns; s++) {
            o_stream r;
            stream_init(&r, seed, STREAM_MULT, (uint32_t)s, (uint32_t)(ev_begin + ev), 0);
            int64_t n;
            if (model == 1) {
                const double u = stream_next(&r);
                n = (int64_t)lambda[s];
                if (u < lambda[s] - (double)n) n++;
            } else if (model == 10 || model == 20) {
                const int64_t dN_int = (int64_t)lambda[s];
                const double k = (model == 10) ? para1*(lambda[s] - (double)dN_int) : para1*lambda[s];
                if (k < 1e-15) n = dN_int;
                else n = ((model == 10) ? dN_int : 0) + nbd_draw(&r, k, para1);
            } else {
                n = poisson_mode(lambda[s], pmode[s], stream_next(&r));
            }
            int64_t nout = n;
            if (lcc == 1) { /* src/FSSW.cpp:931-938, 1035-1048 */
                if (sp[s].charge < 0) { n = 0; nout = 0; }
                else if (sp[s].charge > 0) nout = 2*n;
            }
            mult[ev*ns + s] = n;
            out_count[ev*ns + s] = nout;
        }
}

/* ------------------------------------------------------------------ momentum tables */
typedef struct {
    int n, trunc, fermion;
    double m0;
    double *E, *c0, *c1, *c2;
} o_mtab;

static double cdf0(const o_mtab *t, double Et) {
    const double mt = t->m0;
    if (t->trunc > 5) {
        if (t->fermion) return -exp(mt)*log((1. + exp(-Et))/(1. + exp(-mt)));
        return exp(mt)*log((1. - exp(-Et))/(1. - exp(-mt)));
    }
    double res = 0.;
    int sign = 1;
    for (int n = 0; n < t->trunc; n++) {
        const int n1 = n + 1;
        if (t->fermion) {
            res += (sign/n1)*exp(-mt*n)*(1. - exp((mt - Et)*n1)); /* integer division, as the reference */
            sign *= -1;
        } else {
            res += (1./n1)*exp(-mt*n)*(1. - exp((mt - Et)*n1));
        }
    }
    return res;
}

static double cdf1(const o_mtab *t, double Et) {
    const double mt = t->m0;
    double res = 0., sign = 1.;
    for (int n = 0; n < t->trunc; n++) {
        const int n1 = n + 1;
        res += (sign/(n1*n1)*exp(-mt*n)*((mt*n1 + 1) - exp((mt - Et)*n1)*(Et*n1 + 1)));
        if (t->fermion) sign *= -1.;
    }
    return res;
}

static double cdf2(const o_mtab *t, double Et) {
    const double mt = t->m0;
    double res = 0.;
    for (int n = 0; n < t->trunc; n++) {
        const int n1 = n + 1;
        res += (1./(n1*n1*n1)*exp(-mt*n)*((mt*n1*(mt*n1 + 2) + 2)
                                           - exp((mt - Et)*n1)*(Et*n1*(Et*n1 + 2) + 2)));
    }
    return res;
}

static void mtab_build(o_mtab *t, int fermion, double m0, int trunc) {
    t->fermion = fermion;
    t->m0 = m0;
    t->trunc = trunc;
    double Emin, Emax, dE;
    if (fermion) { Emin = m0; Emax = Emin + 40.; dE = 0.02; }
    else { Emin = m0 + 0.05; Emax = Emin + 50.; dE = 0.05; }
    const int n = (Emax - Emin)/dE + 1;
    t->n = n;
    t->E = malloc(sizeof(double)*4*n);
    t->c0 = t->E + n; t->c1 = t->c0 + n; t->c2 = t->c1 + n;
    for (int i = 0; i < n; i++) {
        const double Et = Emin + i*dE;
        t->E[i] = Et;
        t->c0[i] = cdf0(t, Et);
        t->c1[i] = cdf1(t, Et);
        t->c2[i] = cdf2(t, Et);
    }
}

static o_mtab g_tab[6]; /* boson regimes 0..2, fermion regimes 0..2 */
static int g_tab_ready = 0;

static void tables_ready(void) {
    if (g_tab_ready) return;
    const double m0b[3] = {0.05, 30., 50.}, m0f[3] = {0., 30., 50.};
    const int tr[3] = {10, 2, 1};
    for (int r = 0; r < 3; r++) {
        mtab_build(&g_tab[r], 0, m0b[r], tr[r]);
        mtab_build(&g_tab[3 + r], 1, m0f[r], tr[r]);
    }
    g_tab_ready = 1;
}

/* table r as [4][n] (Etilde, CDF_0, CDF_1, CDF_2); returns n */
int oracle_momentum_table(int r, double *dst) {
    tables_ready();
    if (dst) memcpy(dst, g_tab[r].E, sizeof(double)*4*g_tab[r].n);
    return g_tab[r].n;
}

typedef struct {
    const o_mtab *t;
    double T, mu, m_tilde, mu_tilde, m_term, cdf_max, a;
    int idx_min;
} o_psetup;

static double tabF(const o_psetup *q, int i) {
    const o_mtab *t = q->t;
    return (t->c2[i] + 2.*q->mu_tilde*t->c1[i]
            + (q->mu_tilde*q->mu_tilde - q->m_tilde*q->m_tilde/2.)*t->c0[i] - q->m_term);
}

/* returns 0 when (m - mu)/T is outside the table (the reference exits, MomentumSamplerBase.cpp:35-43) */
static int psetup(o_psetup *q, double m, double T, double mu, int sign) {
    const double m0tilde = m/T - mu/T;
    const int regime = (m0tilde < 30.) ? 0 : (m0tilde < 50. ? 1 : 2);
    q->t = &g_tab[(sign == -1 ? 0 : 3) + regime];
    T = fmax(1e-16, T);
    q->T = T;
    q->mu = mu;
    q->m_tilde = m/T;
    q->mu_tilde = mu/T;
    q->a = q->m_tilde - q->mu_tilde;
    const double w0 = q->mu_tilde*q->mu_tilde - q->m_tilde*q->m_tilde/2.;
    q->m_term = cdf2(q->t, q->a) + 2.*q->mu_tilde*cdf1(q->t, q->a) + w0*cdf0(q->t, q->a);
    const int idx_max = q->t->n - 1;
    q->cdf_max = tabF(q, idx_max);
    q->idx_min = (int)((q->a - q->t->E[0])/(q->t->E[1] - q->t->E[0]));
    return !(q->idx_min < 0 || q->idx_min >= idx_max);
}

static double inverse_cdf(const o_psetup *q, double r) {
    int lo = q->idx_min, hi = q->t->n - 1;
    double r_min = tabF(q, lo), r_max = q->cdf_max;
    while (hi - lo > 1) {
        const int mid = (int)((hi + lo)/2 + 0.1);
        const double r_mid = tabF(q, mid);
        if (r < r_mid) { hi = mid; r_max = r_mid; }
        else { lo = mid; r_min = r_mid; }
    }
    double E0 = q->t->E[lo];
    if (E0 < q->a) { E0 = q->a; r_min = 0.; }
    return E0 + (q->t->E[hi] - E0)/fmax(1e-16, r_max - r_min)*(r - r_min);
}

/* MomentumSamplerBase::Sample_a_momentum's do-while; one block per iteration.  The last word of
 * the accepted iteration's block is returned for the azimuth. */
static double sample_p(const o_psetup *q, double m, o_stream *rng, uint32_t *phi_word) {
    double p, E, ratio;
    uint32_t w[4];
    do {
        stream_block(rng, w);
        const double r = w53(w[0], w[1])*q->cdf_max;
        const double Et = inverse_cdf(q, r);
        E = q->T*Et + q->mu;
        p = sqrt(E*E - m*m);
        ratio = (p/E)/(1. - m*m/(2.*E*E));
    } while (w32(w[2]) > ratio);
    *phi_word = w[3];
    return p;
}

/* n |p| samples for fixed (m, T, mu, sign): unit test of rows M of SURVEY.md section 8(a) */
int oracle_sample_momentum(double m, double T, double mu, int sign, int64_t n, uint64_t seed,
                           double *out) {
    tables_ready();
    o_psetup q;
    if (!psetup(&q, m, T, mu, sign)) return 1;
    for (int64_t i = 0; i < n; i++) {
        o_stream rng;
        stream_init(&rng, seed, STREAM_SAMPLE, 0, (uint32_t)(i >> 20), (uint32_t)(i & 0xFFFFF));
        uint32_t unused;
        out[i] = sample_p(&q, m, &rng, &unused);
    }
    return 0;
}

/* ------------------------------------------------------------------ one hadron */
/* largest i with cdf[i] < v, cdf[0] = 0 (RandomVariable1DArray::rand + binarySearch) */
static int64_t pick_cell(const double *cdf, int64_t ncell, double u) {
    const double v = (cdf[ncell] - 1e-15)*u;
    int64_t lo = 0, hi = ncell;
    while (hi - lo > 1) {
        const int64_t mid = (lo + hi)/2;
        if (cdf[mid] < v) lo = mid; else hi = mid;
    }
    return lo;
}

static double clip01(double x) { return fmax(0., fmin(1., x)); }

/* FSSW::sample_momemtum_from_a_fluid_cell: returns 1 on accept, 0 after 4999 rejected tries */
static int sample_in_cell(const float *c, const double *coef, const o_options *o, double mass,
                          int sign, int B, int S, int Q, o_stream *rng, int *ntries, double *pT,
                          double *phi, double *y_minus_eta, int *range_error) {
    const double Tdec = c[F_T];
    const float muf = B*c[F_MUB] + S*c[F_MUS] + Q*c[F_MUQ];
    const double mu = fmin(mass, (double)muf);
    const float e_plus_p = c[F_E] + c[F_P];
    const double deltaf_prefactor = 1.0/(2.0*Tdec*Tdec*e_plus_p);
    const double prefactor_qmu = c[F_NB]/e_plus_p;
    const float d2 = c[F_DA1]*c[F_DA1] + c[F_DA2]*c[F_DA2] + c[F_DA3]*c[F_DA3];
    const double dsigma_fac = fabsf(c[F_DA0]) + sqrt((double)d2);
    const int neos = (o->bulk_kind == 21) ? 1 : (o->bulk_kind == 20 ? 0 : -1);
    o_psetup q;
    if (!psetup(&q, mass, Tdec, mu, sign)) { *range_error = 1; return 0; }
    int tries = 1;
    while (tries < 5000) {
        uint32_t phi_word, w[4];
        const double p_mag = sample_p(&q, mass, rng, &phi_word);
        (*ntries)++;
        stream_block(rng, w);
        const double phi_c = 2*M_PI*w32(phi_word);
        const double cos_theta = 2.*w32(w[0]) - 1.;
        const double sin_theta = sqrt(1. - cos_theta*cos_theta);
        const double pT_c = p_mag*sin_theta;
        const double px = pT_c*cos(phi_c), py = pT_c*sin(phi_c);
        const double p0 = sqrt(mass*mass + p_mag*p_mag);
        const double pz = p_mag*cos_theta;
        const double pdsigma = p0*c[F_DA0] + px*c[F_DA1] + py*c[F_DA2] + pz*c[F_DA3];
        const double f0 = 1./(exp((p0 - mu)/Tdec) + sign);
        double d_shear = 0., d_bulk = 0., d_q = 0.;
        if (o->include_shear == 1) {
            const double W = (px*px*c[F_PIXX] + 2.*px*py*c[F_PIXY] + 2.*px*pz*c[F_PIXZ]
                              + py*py*c[F_PIYY] + 2.*py*pz*c[F_PIYZ]
                              + pz*pz*(-c[F_PIXX] - c[F_PIYY]));
            if (neos == 1) d_shear = (1. - sign*f0)*W/(2.*coef[2])/(p0*Tdec);
            else if (neos == 0) d_shear = (1. - sign*f0)*W*coef[0];
            else d_shear = (1. - sign*f0)*W*deltaf_prefactor;
        }
        if (o->include_bulk == 1) {
            double bulkPi = 0.;
            const int k = o->bulk_kind;
            if (k == 11 || k == 21 || k == 20) bulkPi = c[F_BULK];
            else if (k == 1) bulkPi = c[F_BULK]/HBARC;
            if (k == 1 || k == 21) {
                const double EoT = p0/Tdec, moT = mass/Tdec;
                d_bulk = (-1.0*(1. - sign*f0)*coef[0]*(moT*moT/(3.*EoT) - coef[1]*EoT)*bulkPi);
            } else if (k == 11) {
                d_bulk = ((1. - sign*f0)*bulkPi*(coef[0]*mass*mass + coef[1]*B*p0 + coef[2]*p0*p0));
            } else if (k == 20) {
                d_bulk = ((1. - sign*f0)*bulkPi*(mass*mass*coef[2]
                                                 + p0*(B*coef[3] + S*coef[4] + Q*coef[5])
                                                 + p0*p0*(coef[1] - coef[2])));
            } else if (k == 0) {
                /* bulkPi stays 0 for kinds 0,2,3,4 in FSSW (src/FSSW.cpp:1918-1924) */
                d_bulk = 0.;
            }
        }
        if (o->include_diff == 1) {
            const double qf = (-px*c[F_QX] - py*c[F_QY] - pz*c[F_QZ]);
            d_q = ((1. - sign*f0)*(prefactor_qmu - B/p0)*qf/coef[6]);
        }
        const double fact1 = clip01(pdsigma/p0/dsigma_fac);
        const double fact2 = clip01((1. + d_shear + d_bulk + d_q)/2.);
        if (w32(w[1]) < fact1*fact2) {
            const float pl[4] = {(float)p0, (float)px, (float)py, (float)pz};
            const float *u = c + F_UT;
            double pdu = 0.;
            for (int i = 1; i < 4; i++) pdu += pl[i]*u[i];
            float lab[4];
            lab[0] = pl[0]*u[0] + pdu;
            for (int i = 1; i < 4; i++) lab[i] = pl[i] + (pdu/(u[0] + 1) + pl[0])*u[i];
            *pT = sqrt(lab[1]*lab[1] + lab[2]*lab[2]);
            *phi = atan2(lab[2], lab[1]);
            const double mT = sqrt((*pT)*(*pT) + mass*mass);
            *y_minus_eta = asinh(lab[3]/mT) - c[F_ETA];
            return 1;
        }
        tries++;
    }
    return 0;
}

/* FSSW::add_one_sampled_particle */
static void emit(o_hadron *h, int pid, double mass, const float *c, double pT, double phi,
                 double y_minus_eta, double eta_s) {
    const double y = y_minus_eta + eta_s;
    const double mT = sqrt(mass*mass + pT*pT);
    h->pid = pid;
    h->mass = mass;
    h->E = mT*cosh(y);
    h->px = pT*cos(phi);
    h->py = pT*sin(phi);
    h->pz = mT*sinh(y);
    h->t = c[F_TAU]*cosh(eta_s);
    h->x = c[F_X];
    h->y = c[F_Y];
    h->z = c[F_TAU]*sinh(eta_s);
}

/*
 * The event/species/particle loops of FSSW::sample_using_dN_dxtdy_4all_particles_conventional
 * (src/FSSW.cpp:920-1050) with the engine's stream keying: hadron k of species s in event ev draws
 * from stream (seed; SAMPLE, s, ev, k) in the order (53 = 53-bit uniform from two Philox words,
 * 32 = 32-bit uniform from one word)
 *    cell | per try: (proposal block)+ , direction/accept block | after 4999 rejections: new cell
 *    (one Philox block per decision point, see stream_block above)
 *    | boost-invariant: rapidity | charge-conservation partner: tries in the same cell, same eta_s.
 * Output: event-major, species in sampling order inside an event (as Hadron_list), draw order.
 * yields: [ns][ncell] (cell CDF = sequential exclusive prefix, RandomVariable1DArray.cpp:38-50).
 * Returns the number of hadrons written, or -1 (capacity) / -2 (momentum table range).
 */
int64_t oracle_sample(const float *cells, int64_t ncell, const double *coef /*[ncell][7]*/,
                      const double *yields, const o_species *sp, int ns, const o_options *o,
                      uint64_t seed, int64_t ev_begin, int64_t nev, const int64_t *mult /*[nev][ns]*/,
                      o_hadron *out, int64_t cap, int32_t *out_cell, int32_t *out_tries) {
    tables_ready();
    /* CDFs of all species */
    double *cdf = malloc(sizeof(double)*(size_t)ns*(ncell + 1));
    for (int s = 0; s < ns; s++) {
        double *c = cdf + (size_t)s*(ncell + 1);
        c[0] = 0.;
        for (int64_t l = 0; l < ncell; l++) c[l + 1] = c[l] + fmax(yields[(size_t)s*ncell + l], 0.);
    }
    int64_t n = 0;
    int range_error = 0;
    for (int64_t ev = 0; ev < nev && !range_error; ev++)
        for (int s = 0; s < ns && !range_error; s++) {
            const o_species *p = &sp[s];
            const int64_t N = mult[ev*ns + s];
            for (int64_t k = 0; k < N; k++) {
                o_stream rng;
                stream_init(&rng, seed, STREAM_SAMPLE, (uint32_t)s, (uint32_t)(ev_begin + ev), (uint32_t)k);
                int ntries = 0;
                double pT, phi, yme;
                int64_t cell;
                for (;;) {
                    uint32_t w[4];
                    stream_block(&rng, w);
                    cell = pick_cell(cdf + (size_t)s*(ncell + 1), ncell, w53(w[0], w[1]));
                    if (sample_in_cell(cells + cell*NFIELD, coef + cell*7, o, p->mass, p->sign,
                                       p->baryon, p->strange, p->charge, &rng, &ntries, &pT, &phi,
                                       &yme, &range_error))
                        break;
                    if (range_error) break;
                }
                if (range_error) break;
                const float *c = cells + cell*NFIELD;
                double eta_s = c[F_ETA];
                if (o->hydro_mode != 2) {
                    uint32_t w[4];
                    stream_block(&rng, w);
                    const double rap = o->y_LB + (o->y_RB - o->y_LB)*w32(w[0]);
                    eta_s = rap - yme;
                }
                if (n >= cap) { free(cdf); return -1; }
                emit(&out[n], p->pid, p->mass, c, pT, phi, yme, eta_s);
                if (out_cell) { out_cell[n] = (int32_t)cell; out_tries[n] = ntries; }
                n++;
                if (o->lcc == 1 && p->charge > 0) {
                    int ok, nt2 = 0;
                    do {
                        ok = sample_in_cell(c, coef + cell*7, o, p->mass, p->sign, -p->baryon,
                                            -p->strange, -p->charge, &rng, &nt2, &pT, &phi, &yme,
                                            &range_error);
                    } while (!ok && !range_error);
                    if (range_error) break;
                    /* the partner keeps the primary's eta_s (src/FSSW.cpp:1045-1047) */
                    if (n >= cap) { free(cdf); return -1; }
                    emit(&out[n], -p->pid, p->mass, c, pT, phi, yme, eta_s);
                    if (out_cell) { out_cell[n] = (int32_t)cell; out_tries[n] = nt2; }
                    n++;
                }
            }
        }
    free(cdf);
    return range_error ? -2 : n;
}

/* ------------------------------------------------------------------ decays */
typedef struct {
    int idx;
    float mass, E, px, py, pz, t, x, y, z;
} o_part;

static void boost_d(o_part *d, double E, double px, double py, double pz, double vx, double vy,
                    double vz, double v2, double gamma) {
    const double gm1 = gamma - 1.;
    const double vp = vx*px + vy*py + vz*pz;
    d->E = gamma*(E + vp);
    d->px = px + (gm1*vp/v2 + gamma*E)*vx;
    d->py = py + (gm1*vp/v2 + gamma*E)*vy;
    d->pz = pz + (gm1*vp/v2 + gamma*E)*vz;
}

static int find_pid(const o_dspecies *sp, int nsp, int pid) {
    for (int i = 0; i < nsp; i++)
        if (sp[i].pid == pid) return i;
    return -1;
}

/* one decay of mother m (particle_decay::perform_decays): daughters appended to d[], returns count
 * (0 for channels that are neither 2- nor 3-body), -1 on kinematic/table errors */
static int decay_once(const o_part *m, const o_dspecies *sp, const o_dchannel *ch, o_stream *rng,
                      o_part *d) {
    const o_dspecies *ms = &sp[m->idx];
    const double u = stream_next(rng);
    double cum = 0.;
    int pick = ms->first_channel + ms->n_channels - 1;
    for (int c = 0; c < ms->n_channels; c++) {
        cum += ch[ms->first_channel + c].br;
        if (cum > u) { pick = ms->first_channel + c; break; }
    }
    const o_dchannel *pc = &ch[pick];
    if (pc->n_part != 2 && pc->n_part != 3) return 0;
    for (int i = 0; i < pc->n_part; i++)
        if (pc->daughter[i] < 0) return -1;
    const double M = m->mass, width = ms->width;
    const double vx = m->px/m->E, vy = m->py/m->E, vz = m->pz/m->E;
    const double v2 = vx*vx + vy*vy + vz*vz;
    const double gamma = 1./sqrt(1. - v2);
    for (int i = 0; i < pc->n_part; i++) {
        d[i].idx = pc->daughter[i];
        d[i].mass = sp[pc->daughter[i]].mass;
    }
    if (pc->n_part == 2) {
        const double m1 = d[0].mass, m2 = d[1].mass;
        if (M < m1 + m2) return -1;
        const double temp = M*M - m1*m1 - m2*m2;
        const double p = sqrt(temp*temp - 4*m1*m1*m2*m2)/(2*M);
        const double phi = stream_next(rng)*2*M_PI;
        const double ct = 2.*(stream_next(rng) - 0.5);
        const double st = sqrt(1. - ct*ct);
        const double E1 = sqrt(p*p + m1*m1), E2 = sqrt(p*p + m2*m2);
        const double px = p*st*cos(phi), py = p*st*sin(phi), pz = p*ct;
        boost_d(&d[0], E1, px, py, pz, vx, vy, vz, v2, gamma);
        boost_d(&d[1], E2, -px, -py, -pz, vx, vy, vz, v2, gamma);
        double life = 1e10;
        if (width > 1e-10) {
            const double tau0 = m->E/(M)*1./(width);
            life = -tau0*log(stream_next(rng))*0.19733;
        }
        for (int i = 0; i < 2; i++) {
            d[i].t = m->t + life;
            d[i].x = m->x + m->px/m->E*life;
            d[i].y = m->y + m->py/m->E*life;
            d[i].z = m->z + m->pz/m->E*life;
        }
        return 2;
    }
    const double m1 = d[0].mass, m2 = d[1].mass, m3 = d[2].mass;
    if (M < m1 + m2 + m3) return -1;
    double E1, E2, E3, p1, p2, c12;
    do {
        do {
            E1 = stream_next(rng)*(M - m1 - m2 - m3) + m1;
            E2 = stream_next(rng)*(M - m1 - m2 - m3) + m2;
        } while (E1 + E2 > M);
        p1 = sqrt(E1*E1 - m1*m1);
        p2 = sqrt(E2*E2 - m2*m2);
        E3 = M - E1 - E2;
        c12 = (E3*E3 - p1*p1 - p2*p2 - m3*m3)/(2.*p1*p2);
    } while (c12 < -1.0 || c12 > 1.0);
    double life = 1e10;
    if (width > 1e-10) {
        const double tau = m->E/(M)*1./width;
        life = -tau*log(stream_next(rng))*0.19733;
    }
    const double dt = m->t + life, dx = m->x + m->px/m->E*life, dy = m->y + m->py/m->E*life,
                 dz = m->z + m->pz/m->E*life;
    const double t2x = p2*sqrt(1. - c12*c12), t2z = p2*c12, t3x = -t2x, t3z = -(p1 + t2z);
    const double phi = 2.*M_PI*stream_next(rng), ksi = 2.*M_PI*stream_next(rng);
    const double ct = 2.*stream_next(rng) - 1.0;
    const double sphi = sin(phi), cphi = cos(phi), sksi = sin(ksi), cksi = cos(ksi);
    const double st = sqrt(1. - ct*ct);
    const double p1x = -p1*st*cksi, p1y = p1*st*sksi, p1z = p1*ct;
    E1 = sqrt(m1*m1 + p1x*p1x + p1y*p1y + p1z*p1z);
    const double p2x = (t2x*(cphi*ct*cksi - sphi*sksi) - t2z*st*cksi);
    const double p2y = (t2x*(-cphi*ct*sksi - sphi*cksi) + t2z*st*sksi);
    const double p2z = t2x*(cphi*st) + t2z*ct;
    E2 = sqrt(m2*m2 + p2x*p2x + p2y*p2y + p2z*p2z);
    const double p3x = (t3x*(cphi*ct*cksi - sphi*sksi) - t3z*st*cksi);
    const double p3y = (t3x*(-cphi*ct*sksi - sphi*cksi) + t3z*(st*sksi));
    const double p3z = t3x*cphi*st + t3z*ct;
    E3 = sqrt(m3*m3 + p3x*p3x + p3y*p3y + p3z*p3z);
    boost_d(&d[0], E1, p1x, p1y, p1z, vx, vy, vz, v2, gamma);
    boost_d(&d[1], E2, p2x, p2y, p2z, vx, vy, vz, v2, gamma);
    boost_d(&d[2], E3, p3x, p3y, p3z, vx, vy, vz, v2, gamma);
    for (int i = 0; i < 3; i++) { d[i].t = dt; d[i].x = dx; d[i].y = dy; d[i].z = dz; }
    return 3;
}

/*
 * FSSW::perform_resonance_feed_down with the engine's ordering and keying: primary k of event ev
 * owns stream (seed; DECAY, 0, ev, k); its decay tree is walked depth-first (first daughter first)
 * and its stable descendants are written in that order in place of the primary.  The reference
 * appends unstable daughters to the end of a work list instead (src/FSSW.cpp:1761-1776): same
 * multiset per event, different order, different RNG consumption.
 * Returns the number of hadrons written, -1 capacity, -2 table/kinematics error, -3 stack overflow.
 */
int64_t oracle_decay(const o_hadron *in, const int64_t *event_off, int64_t nev, int64_t ev_begin,
                     const o_dspecies *sp, int nsp, const o_dchannel *ch, uint64_t seed,
                     o_hadron *out, int64_t cap, int64_t *event_off_out) {
    int64_t n = 0;
    o_part stack[64];
    for (int64_t ev = 0; ev < nev; ev++) {
        event_off_out[ev] = n;
        for (int64_t i = event_off[ev]; i < event_off[ev + 1]; i++) {
            const o_hadron *h = &in[i];
            o_stream rng;
            stream_init(&rng, seed, STREAM_DECAY, 0, (uint32_t)(ev_begin + ev), (uint32_t)(i - event_off[ev]));
            const int idx = find_pid(sp, nsp, h->pid);
            if (idx < 0) {
                if (n >= cap) return -1;
                out[n++] = *h;
                continue;
            }
            int top = 0;
            o_part p = {idx, h->mass, h->E, h->px, h->py, h->pz, h->t, h->x, h->y, h->z};
            stack[top++] = p;
            while (top > 0) {
                const o_part m = stack[--top];
                if (sp[m.idx].stable == 1) {
                    if (n >= cap) return -1;
                    o_hadron o = {sp[m.idx].pid, m.mass, m.E, m.px, m.py, m.pz, m.t, m.x, m.y, m.z};
                    out[n++] = o;
                    continue;
                }
                o_part d[3];
                const int nd = decay_once(&m, sp, ch, &rng, d);
                if (nd < 0) return -2;
                if (top + nd > 64) return -3;
                for (int j = nd - 1; j >= 0; j--) stack[top++] = d[j];
            }
        }
    }
    event_off_out[nev] = n;
    return n;
}

/* single decays of identical mothers: the analogue of `ref_driver decay` for the golden test */
int64_t oracle_decay_once_many(int pid, int64_t n, uint64_t seed, const o_dspecies *sp, int nsp,
                               const o_dchannel *ch, const o_hadron *mother, o_hadron *out /*[3n]*/,
                               int32_t *nd_out /*[n]*/) {
    const int idx = find_pid(sp, nsp, pid);
    if (idx < 0) return -2;
    int64_t w = 0;
    for (int64_t i = 0; i < n; i++) {
        o_stream rng;
        stream_init(&rng, seed, STREAM_DECAY, 0, (uint32_t)(i >> 20), (uint32_t)(i & 0xFFFFF));
        o_part m = {idx, mother->mass, mother->E, mother->px, mother->py, mother->pz, mother->t,
                    mother->x, mother->y, mother->z};
        o_part d[3];
        const int nd = decay_once(&m, sp, ch, &rng, d);
        if (nd < 0) return -2;
        nd_out[i] = nd;
        for (int j = 0; j < nd; j++) {
            o_hadron o = {sp[d[j].idx].pid, d[j].mass, d[j].E, d[j].px, d[j].py, d[j].pz, d[j].t,
                          d[j].x, d[j].y, d[j].z};
            out[w++] = o;
        }
    }
    return w;
}

/* ================================================================================================
 * Legacy "conventional" sampler of EmissionFunctionArray (MC_sampling = 2):
 * sample_using_dN_dxtdy_4all_particles_conventional (src/emissionfunction.cpp:3273-3623),
 * estimate_maximum (:4006-4153, 4309-4421), sample_momemtum_from_a_fluid_cell (:4188-4306),
 * get_deltaf_bulk (:4156-4186), add_one_sampled_particle (:4423-4475).
 * Cells are the LAB-frame (Milne) records, 32 floats in the ISS_L_* order of include/iss_cuda.h,
 * plus pos[4] = x, y, eta_s, 0.  coef[4] per cell = bulk coefficients c0, c1, c2
 * (getbulkvisCoefficients, :3625-3762) and kappa_hat (get_deltaf_qmu_coeff, :3788-3823), which are
 * species independent and restated in numpy (oracle/legacy_oracle.py).
 * ================================================================================================ */
enum { L_TAU = 0, L_U0, L_U1, L_U2, L_U3, L_DA0, L_DA1, L_DA2, L_DA3, L_T, L_P, L_E, L_MUB, L_MUS,
       L_MUQ, L_PI00, L_PI01, L_PI02, L_PI03, L_PI11, L_PI12, L_PI13, L_PI22, L_PI23, L_PI33,
       L_BULK, L_BN, L_Q0, L_Q1, L_Q2, L_Q3, L_SPARE, L_NFIELD };

typedef struct {
    int32_t include_shear, include_bulk, bulk_kind, include_diff;
    int32_t restrict_deltaf, boost_invariant, lcc, reserved1;
    double deltaf_max_ratio, pT_to, y_minus_eta_s_range, y_LB, y_RB;
} o_legacy_opt;

/* principal branch of the Lambert W function for x >= 0 (Halley iteration); the reference calls
 * gsl_sf_lambert_W0 (third party, not in the tree) */
static double lambert_w0(double x) {
    if (x == 0.) return 0.;
    double w = (x < 1.) ? x*(1. - x + 1.5*x*x) : log(x) - log(log(x) + 1.);
    if (!(w > 0.)) w = 0.5;
    for (int it = 0; it < 60; it++) {
        const double e = exp(w), f = w*e - x;
        const double dw = f/(e*(w + 1.) - (w + 2.)*f/(2.*w + 2.));
        w -= dw;
        if (fabs(dw) <= 1e-16*(1. + fabs(w))) break;
    }
    return w;
}

#define LAMBERT_N 40001     /* (200 - 0)/0.005 + 1, :3858-3870 */
static double lambert_tab[LAMBERT_N];
static int lambert_ready = 0;
static double legacy_lambertW(double arg) {     /* get_special_function_lambertW, :3936-3953 */
    const double x_min = 0., x_max = 200.0, dx = 0.005;
    if (!lambert_ready) {
        for (int i = 0; i < LAMBERT_N; i++) lambert_tab[i] = lambert_w0(x_min + i*dx);
        lambert_ready = 1;
    }
    if (arg < x_min || arg > x_max - dx) return lambert_w0(arg);
    const int idx = (int)((arg - x_min)/dx);
    const double fraction = (arg - x_min - idx*dx)/dx;
    return (1. - fraction)*lambert_tab[idx] + fraction*lambert_tab[idx + 1];
}
double oracle_lambert_w0(double x) { return lambert_w0(x); }

/* TableFunction::map with interpolation_model 5 = interpCubicDirect without extrapolation
 * (src/arsenal.cpp:58-110) on iSS_tables/z_exp_m_z.dat; the reference exit(1)s outside the table,
 * here NaN is returned */
static double z_map(const double *x, const double *y, int size, double xx) {
    const double dx = x[1] - x[0];
    if (fabs(xx - x[0]) < dx*1e-30) return y[0];
    const long idx = (long)floor((xx - x[0])/dx);
    if (idx < 0 || idx >= size - 1) return NAN;
    if (idx == 0) {
        const double A0 = y[0], A1 = y[1], A2 = y[2], d = xx - x[0];
        return (A0 - 2.0*A1 + A2)/(2.0*dx*dx)*d*d - (3.0*A0 - 4.0*A1 + A2)/(2.0*dx)*d + A0;
    } else if (idx == size - 2) {
        const double A0 = y[size - 3], A1 = y[size - 2], A2 = y[size - 1];
        const double d = xx - (x[0] + (idx - 1)*dx);
        return (A0 - 2.0*A1 + A2)/(2.0*dx*dx)*d*d - (3.0*A0 - 4.0*A1 + A2)/(2.0*dx)*d + A0;
    }
    const double A0 = y[idx - 1], A1 = y[idx], A2 = y[idx + 1], A3 = y[idx + 2];
    const double d = xx - (x[0] + idx*dx);
    return (-A0 + 3.0*A1 - 3.0*A2 + A3)/(6.0*dx*dx*dx)*d*d*d + (A0 - 2.0*A1 + A2)/(2.0*dx*dx)*d*d
           - (2.0*A0 + 3.0*A1 - 6.0*A2 + A3)/(6.0*dx)*d + A1;
}

/* max over E >= mass of E^A f0(E): the common part of estimate_ideal_maximum (A = 1),
 * estimate_shear_viscous_maximum (A = 3) and estimate_diffusion_maximum (A = 2) */
static double legacy_power_max(int A, int sign, double mass, double T, double mu, double f0_mass,
                               const double *zx, const double *zy, int nz) {
    const double inv_T = 1./T;
    double massA = mass;
    for (int i = 1; i < A; i++) massA *= mass;
    if (sign == 1) {
        double Emax = T*(legacy_lambertW(A*exp(inv_T*mu - A)) + A);
        if (Emax < mass) Emax = mass;
        double EA = Emax;
        for (int i = 1; i < A; i++) EA *= Emax;
        return EA/(exp((Emax - mu)*inv_T) + sign);
    }
    const double rhs = A*exp(inv_T*mu - A);
    if (rhs > 0.3678794) return massA*f0_mass;
    const double Emax = T*(A - z_map(zx, zy, nz, rhs));
    if (Emax < mass) return massA*f0_mass;      /* also taken when the table look-up is NaN: no */
    double EA = Emax;
    for (int i = 1; i < A; i++) EA *= Emax;
    const double g1 = EA/(exp((Emax - mu)*inv_T) + sign), g2 = massA*f0_mass;
    return g1 > g2 ? g1 : g2;
}

double oracle_legacy_estimate_maximum(const float *c, const double *coef, const o_legacy_opt *o,
                                      double mass, int sign, int degen, int baryon, int strange,
                                      int charge, const double *zx, const double *zy, int nz) {
    const double prefactor = 1.0/(8.0*(M_PI*M_PI*M_PI))/HBARC/HBARC/HBARC;
    const double Tdec = c[L_T], inv_Tdec = 1.0/Tdec, Pdec = c[L_P], Edec = c[L_E];
    const double mu = baryon*c[L_MUB] + strange*c[L_MUS] + charge*c[L_MUQ];  /* float arithmetic */
    double bulkPi = 0.0;
    if (o->include_bulk == 1) bulkPi = (o->bulk_kind == 0) ? (double)c[L_BULK] : c[L_BULK]/HBARC;
    double prefactor_qmu = 0.0;
    if (o->include_diff == 1) prefactor_qmu = (double)c[L_BN]/(Edec + Pdec);
    const double u_dot_dsigma = (double)(c[L_TAU]*(c[L_U0]*c[L_DA0] + c[L_U1]*c[L_DA1] + c[L_U2]*c[L_DA2]
                                                   + c[L_U3]*c[L_DA3]/c[L_TAU]));
    const double dsigma_sq = (double)(c[L_TAU]*c[L_TAU]*(c[L_DA0]*c[L_DA0] - c[L_DA1]*c[L_DA1]
                                      - c[L_DA2]*c[L_DA2] - c[L_DA3]*c[L_DA3]/(c[L_TAU]*c[L_TAU])));
    const double dsigmaT = sqrt(fabs(dsigma_sq - u_dot_dsigma*u_dot_dsigma));
    const double dsigma_all = fabs(u_dot_dsigma) + dsigmaT;
    const double f0_mass = 1./(exp((mass - mu)*inv_Tdec) + sign);
    const double guess_ideal = legacy_power_max(1, sign, mass, Tdec, mu, f0_mass, zx, zy, nz);
    double guess_viscous = 0.0;
    if (o->include_shear == 1) {
        const double pi[10] = {c[L_PI00], c[L_PI01], c[L_PI02], c[L_PI03], c[L_PI11], c[L_PI12],
                               c[L_PI13], c[L_PI22], c[L_PI23], c[L_PI33]};
        const double trace_Pi2 = (pi[0]*pi[0] + pi[4]*pi[4] + pi[7]*pi[7] + pi[9]*pi[9]
                                  - 2.*pi[1]*pi[1] - 2.*pi[2]*pi[2] - 2.*pi[3]*pi[3]
                                  + 2.*pi[5]*pi[5] + 2.*pi[6]*pi[6] + 2.*pi[8]*pi[8]);
        const double pi_size = sqrt(trace_Pi2)/(Edec + Pdec);
        const double tmp_factor = (sign == -1) ? 2.0 : 1.0;
        guess_viscous = legacy_power_max(3, sign, mass, Tdec, mu, f0_mass, zx, zy, nz)
                        *(tmp_factor/(2.0*Tdec*Tdec)*pi_size);
    }
    double guess_bulk = 0.0;
    if (o->include_bulk == 1)
        guess_bulk = fabs(bulkPi*coef[0])*mass*mass*inv_Tdec/3.*f0_mass*(1. - sign*f0_mass);
    double guess_qmu = 0.0;
    if (o->include_diff == 1) {
        const float q[4] = {c[L_Q0], c[L_Q1], c[L_Q2], c[L_Q3]};         /* Vec4 is float */
        const double qmu_sq = q[0]*q[0] - q[1]*q[1] - q[2]*q[2] - q[3]*q[3];
        const double q_size = sqrt(fabs(qmu_sq))/coef[3];
        guess_qmu = prefactor_qmu*legacy_power_max(2, sign, mass, Tdec, mu, f0_mass, zx, zy, nz);
        if (baryon > 0) guess_qmu += baryon*guess_ideal;
        guess_qmu *= ((sign == -1) ? 2.0 : 1.0)*q_size;
    }
    return prefactor*degen*dsigma_all*(guess_ideal + guess_viscous + guess_bulk + guess_qmu);
}

static double legacy_deltaf_bulk(const o_legacy_opt *o, double mass, double pdotu, double bulkPi,
                                 double Tdec, int sign, double f0, const double *bc) {
    if (o->include_bulk == 0) return 0.0;       /* get_deltaf_bulk, :4156-4186 */
    const double stat = 1. - sign*f0;
    if (o->bulk_kind == 0) {
        return -stat*bulkPi*(bc[0]*mass*mass + bc[1]*pdotu + bc[2]*pdotu*pdotu);
    } else if (o->bulk_kind == 1) {
        const double E_over_T = pdotu/Tdec, mass_over_T = mass/Tdec;
        return -1.0*stat*bc[0]*(mass_over_T*mass_over_T/(3.*E_over_T) - bc[1]*E_over_T)*bulkPi;
    } else if (o->bulk_kind == 2) {
        const double E_over_T = pdotu/Tdec;
        return -1.*stat*bulkPi*(-bc[0] + bc[1]*E_over_T);
    } else if (o->bulk_kind == 3) {
        const double E_over_T = pdotu/Tdec;
        return -1.*stat*bulkPi/sqrt(E_over_T)*(-bc[0] + bc[1]*E_over_T);
    } else if (o->bulk_kind == 4) {
        const double E_over_T = pdotu/Tdec;
        return -1.*stat*bulkPi*(bc[0] - bc[1]/E_over_T);
    }
    return 0.0;
}

/* EmissionFunctionArray::sample_momemtum_from_a_fluid_cell (:4188-4306): 1 on accept, 0 after
 * 4999 rejected tries.  One Philox block per try: w0 -> pT^2, w1 -> phi, w2 -> y - eta_s,
 * w3 -> accept (32-bit uniforms). */
static int legacy_sample_in_cell(const float *c, const double *coef, const o_legacy_opt *o,
                                 double mass, double degen, int sign, int baryon, int strange,
                                 int charge, double maximum_guess, o_stream *rng, int *ntries,
                                 double *pT_out, double *phi_out, double *yme_out) {
    const double Tdec = c[L_T];
    const double mu = baryon*c[L_MUB] + strange*c[L_MUS] + charge*c[L_MUQ];
    const double inv_Tdec = 1./Tdec;
    const double prefactor = 1.0/(8.0*(M_PI*M_PI*M_PI)*(HBARC*HBARC*HBARC));
    const double deltaf_prefactor = 1.0/(2.0*Tdec*Tdec*(c[L_E] + c[L_P]));
    const double prefactor_qmu = c[L_BN]/(c[L_E] + c[L_P]);
    int tries = 1;
    while (tries < 5000) {
        uint32_t w[4];
        stream_block(rng, w);
        (*ntries)++;
        const double pT = sqrt(o->pT_to*o->pT_to*w32(w[0]));
        const double phi = 2*M_PI*w32(w[1]);
        const double yme = (1. - 2.*w32(w[2]))*o->y_minus_eta_s_range;
        const double mT = sqrt(mass*mass + pT*pT);
        const double px = pT*cos(phi), py = pT*sin(phi);
        const double p0 = mT*cosh(yme), p3 = mT*sinh(yme);
        const double pdotu = p0*c[L_U0] - px*c[L_U1] - py*c[L_U2] - p3*c[L_U3];
        const double expon = (pdotu - mu)*inv_Tdec;
        const double f0 = 1./(exp(expon) + sign);
        const double pdsigma = p0*c[L_DA0] + px*c[L_DA1] + py*c[L_DA2] + p3*c[L_DA3]/c[L_TAU];
        double delta_f_shear = 0.;
        if (o->include_shear == 1) {
            const double Wfactor = (p0*p0*c[L_PI00] - 2.0*p0*px*c[L_PI01] - 2.0*p0*py*c[L_PI02]
                                    - 2.0*p0*p3*c[L_PI03] + px*px*c[L_PI11] + 2.0*px*py*c[L_PI12]
                                    + 2.0*px*p3*c[L_PI13] + py*py*c[L_PI22] + 2.0*py*p3*c[L_PI23]
                                    + p3*p3*c[L_PI33]);
            delta_f_shear = (1. - sign*f0)*Wfactor*deltaf_prefactor;
        }
        double delta_f_bulk = 0.;
        if (o->include_bulk == 1)
            delta_f_bulk = legacy_deltaf_bulk(o, mass, pdotu, c[L_BULK]/HBARC, Tdec, sign, f0, coef);
        double delta_f_qmu = 0.0;
        if (o->include_diff == 1) {
            const double qmufactor = p0*c[L_Q0] - px*c[L_Q1] - py*c[L_Q2] - p3*c[L_Q3];
            delta_f_qmu = (1. - sign*f0)*(prefactor_qmu - baryon/pdotu)*qmufactor/coef[3];
        }
        double resize_factor = 1.0;
        if (o->restrict_deltaf == 1) {
            const double deltaf_size = fabs(delta_f_shear + delta_f_bulk + delta_f_qmu);
            resize_factor = fmin(1., o->deltaf_max_ratio/(deltaf_size + 1e-10));
        }
        const double result = prefactor*degen*f0*pdsigma*c[L_TAU]
                              *(1. + (delta_f_shear + delta_f_bulk + delta_f_qmu)*resize_factor);
        const double accept_prob = result/(1.0*maximum_guess);
        if (w32(w[3]) < accept_prob) {
            *pT_out = pT; *phi_out = phi; *yme_out = yme;
            return 1;
        }
        tries++;
    }
    return 0;
}

/* event/species/particle loops of sample_using_dN_dxtdy_4all_particles_conventional (:3330-3560)
 * with the engine's stream keying (as oracle_sample above): stream (seed; SAMPLE, s, ev, k),
 * block order  cell | tries (one block each) | after 4999 rejections: new cell | boost-invariant:
 * rapidity | charge-conservation partner: tries in the same cell.  yields [ns][ncell] may be negative (the reference does not clamp them; the CDF does,
 * RandomVariable1DArray.cpp:38-50).  cdf_in (may be NULL): [ns][ncell+1] prefix to use instead of
 * the sequential one (the engine's fixed-order prefix).  max_out (may be NULL): maximum_guess used
 * for every hadron.  Returns the number of hadrons, -1 on capacity, -3 if a maximum is NaN (the
 * reference exit(1)s in the z_exp_m_z look-up). */
int64_t oracle_legacy_sample(const float *lab, const float *pos, int64_t ncell,
                             const double *coef /*[ncell][4]*/, const double *yields,
                             const double *cdf_in, const o_species *sp, int ns,
                             const o_legacy_opt *o, const double *zx, const double *zy, int nz,
                             uint64_t seed, int64_t ev_begin, int64_t nev, const int64_t *mult,
                             o_hadron *out, int64_t cap, int32_t *out_cell, int32_t *out_tries) {
    double *cdf = NULL;
    if (!cdf_in) {
        cdf = malloc(sizeof(double)*(size_t)ns*(ncell + 1));
        for (int s = 0; s < ns; s++) {
            double *c = cdf + (size_t)s*(ncell + 1);
            c[0] = 0.;
            for (int64_t l = 0; l < ncell; l++) c[l + 1] = c[l] + fmax(yields[(size_t)s*ncell + l], 0.);
        }
        cdf_in = cdf;
    }
    int64_t n = 0;
    for (int64_t ev = 0; ev < nev; ev++)
        for (int s = 0; s < ns; s++) {
            const o_species *p = &sp[s];
            const int64_t N = mult[ev*ns + s];
            for (int64_t k = 0; k < N; k++) {
                o_stream rng;
                stream_init(&rng, seed, STREAM_SAMPLE, (uint32_t)s, (uint32_t)(ev_begin + ev), (uint32_t)k);
                int ntries = 0;
                double pT = 0, phi = 0, yme = 0;
                int64_t cell;
                for (;;) {
                    uint32_t w[4];
                    stream_block(&rng, w);
                    cell = pick_cell(cdf_in + (size_t)s*(ncell + 1), ncell, w53(w[0], w[1]));
                    const float *c = lab + cell*L_NFIELD;
                    const double mx = oracle_legacy_estimate_maximum(c, coef + cell*4, o, p->mass,
                        p->sign, p->gspin, p->baryon, p->strange, p->charge, zx, zy, nz);
                    if (isnan(mx)) { free(cdf); return -3; }
                    if (legacy_sample_in_cell(c, coef + cell*4, o, p->mass, (double)p->gspin, p->sign,
                                              p->baryon, p->strange, p->charge, mx, &rng, &ntries,
                                              &pT, &phi, &yme))
                        break;
                }
                const float *c = lab + cell*L_NFIELD;
                double eta_s = pos[cell*4 + 2];
                if (o->boost_invariant) {
                    uint32_t w[4];
                    stream_block(&rng, w);
                    const double rap = o->y_LB + (o->y_RB - o->y_LB)*w32(w[0]);
                    eta_s = rap - yme;
                }
                if (n >= cap) { free(cdf); return -1; }
                {   /* add_one_sampled_particle, :4423-4475 */
                    o_hadron *h = &out[n];
                    const double y = yme + eta_s;
                    const double mT = sqrt(p->mass*p->mass + pT*pT);
                    h->pid = p->pid;
                    h->mass = p->mass;
                    h->E = mT*cosh(y);
                    h->px = pT*cos(phi);
                    h->py = pT*sin(phi);
                    h->pz = mT*sinh(y);
                    h->t = c[L_TAU]*cosh(eta_s);
                    h->x = pos[cell*4 + 0];
                    h->y = pos[cell*4 + 1];
                    h->z = c[L_TAU]*sinh(eta_s);
                }
                if (out_cell) { out_cell[n] = (int32_t)cell; out_tries[n] = ntries; }
                n++;
                if (o->lcc == 1 && p->charge > 0) {
                    /* a negative partner from the same cell, same maximum_guess, same eta_s
                     * (:3517-3546); the do-while never draws a new cell */
                    const double mx = oracle_legacy_estimate_maximum(c, coef + cell*4, o, p->mass,
                        p->sign, p->gspin, p->baryon, p->strange, p->charge, zx, zy, nz);
                    int nt2 = 0;
                    double pT2 = 0, phi2 = 0, yme2 = 0;
                    while (!legacy_sample_in_cell(c, coef + cell*4, o, p->mass, (double)p->gspin,
                                                  p->sign, -p->baryon, -p->strange, -p->charge, mx,
                                                  &rng, &nt2, &pT2, &phi2, &yme2)) {
                    }
                    if (n >= cap) { free(cdf); return -1; }
                    o_hadron *h = &out[n];
                    const double y = yme2 + eta_s;
                    const double mT = sqrt(p->mass*p->mass + pT2*pT2);
                    h->pid = -p->pid;
                    h->mass = p->mass;
                    h->E = mT*cosh(y);
                    h->px = pT2*cos(phi2);
                    h->py = pT2*sin(phi2);
                    h->pz = mT*sinh(y);
                    h->t = c[L_TAU]*cosh(eta_s);
                    h->x = pos[cell*4 + 0];
                    h->y = pos[cell*4 + 1];
                    h->z = c[L_TAU]*sinh(eta_s);
                    if (out_cell) { out_cell[n] = (int32_t)cell; out_tries[n] = nt2; }
                    n++;
                }
            }
        }
    free(cdf);
    return n;
}
