// iss_rng.h -- counter-based RNG and the integer multiplicity draw, shared by the
// device kernels and the host facade.  Replaces RandomUtil::Random (reference
// src/Random.h:11-22, std::mt19937) and gsl_ran_poisson (FSSW.cpp:293-298).
//
// Everything here is plain integer arithmetic or single IEEE-754 double
// operations that are never contracted (explicit __dmul_rn/__dadd_rn/__ddiv_rn on
// the device, -ffp-contract=off on the host), so that the integer bookkeeping is
// bit-identical on CPU and GPU (see oracle/iss_oracle.py for the independent
// numpy restatement the tests compare with).
#ifndef ISS_RNG_H_
#define ISS_RNG_H_

#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define ISS_HD __host__ __device__ __forceinline__
#else
#define ISS_HD inline
#endif

#if defined(__CUDA_ARCH__)
#define ISS_MUL(a, b) __dmul_rn((a), (b))
#define ISS_ADD(a, b) __dadd_rn((a), (b))
#define ISS_SUB(a, b) __dadd_rn((a), -(b))
#define ISS_DIV(a, b) __ddiv_rn((a), (b))
#else
#define ISS_MUL(a, b) ((a)*(b))
#define ISS_ADD(a, b) ((a) + (b))
#define ISS_SUB(a, b) ((a) - (b))
#define ISS_DIV(a, b) ((a)/(b))
#endif

namespace iss {

// random streams (upper byte of counter word 3)
enum { STREAM_MULT = 1, STREAM_SAMPLE = 2, STREAM_DECAY = 3 };

struct Philox {
    uint32_t c[4];
    uint32_t k[2];
};

ISS_HD void philox_mulhilo(uint32_t a, uint32_t b, uint32_t &hi, uint32_t &lo) {
#if defined(__CUDA_ARCH__)
    lo = a*b;
    hi = __umulhi(a, b);
#else
    uint64_t p = static_cast<uint64_t>(a)*static_cast<uint64_t>(b);
    lo = static_cast<uint32_t>(p);
    hi = static_cast<uint32_t>(p >> 32);
#endif
}

// Philox4x32-10 (Salmon et al., SC'11): 10 rounds, Weyl key schedule.
ISS_HD void philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int r = 0; r < 10; r++) {
        uint32_t hi0, lo0, hi1, lo1;
        philox_mulhilo(0xD2511F53u, c0, hi0, lo0);
        philox_mulhilo(0xCD9E8D57u, c2, hi1, lo1);
        uint32_t n0 = hi1 ^ c1 ^ k0;
        uint32_t n1 = lo1;
        uint32_t n2 = hi0 ^ c3 ^ k1;
        uint32_t n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// 53-bit uniform in [0,1) from two 32-bit words.
ISS_HD double u53(uint32_t hi, uint32_t lo) {
    uint64_t v = (static_cast<uint64_t>(hi >> 5) << 26) | static_cast<uint64_t>(lo >> 6);
    return static_cast<double>(v)*(1.0/9007199254740992.0);
}

// 32-bit uniform in (0,1): (x + 1/2) 2^-32.
ISS_HD double u32(uint32_t x) {
    return (static_cast<double>(x) + 0.5)*(1.0/4294967296.0);
}

// A per-object random stream: key = seed, counter = (block, draw, event, species|stream<<24).
// Each Philox block yields four 32-bit words that are handed out in order; next() consumes two
// words for a 53-bit uniform in [0,1), next32() one word for a 32-bit uniform in (0,1).
struct Stream {
    uint32_t ctr[4];
    uint32_t key[2];
    uint32_t buf[4];
    int pos;    // next unused word of buf (4: empty)

    ISS_HD void init(uint64_t seed, uint32_t stream, uint32_t species, uint32_t event,
                     uint32_t draw, uint32_t block0 = 0) {
        key[0] = static_cast<uint32_t>(seed);
        key[1] = static_cast<uint32_t>(seed >> 32);
        ctr[0] = block0;
        ctr[1] = draw;
        ctr[2] = event;
        ctr[3] = (stream << 24) | (species & 0xFFFFFFu);
        pos = 4;
    }
    ISS_HD uint32_t word() {
        if (pos == 4) {
            philox4x32_10(ctr, key, buf);
            ctr[0]++;
            pos = 0;
        }
        const uint32_t w = (pos == 0) ? buf[0] : (pos == 1) ? buf[1] : (pos == 2) ? buf[2] : buf[3];
        pos++;
        return w;
    }
    ISS_HD double next() {
        const uint32_t hi = word();
        const uint32_t lo = word();
        return u53(hi, lo);
    }
    ISS_HD double next32() { return u32(word()); }
    // the state after n calls of next() whose values nobody needs (whole blocks are not generated)
    ISS_HD void skip(int n) {
        int w = 2*n - (4 - pos);        // words beyond the buffered ones
        if (w <= 0) {
            pos += 2*n;
            return;
        }
        ctr[0] += static_cast<uint32_t>(w >> 2);
        pos = 4;
        if (w & 3) {
            philox4x32_10(ctr, key, buf);
            ctr[0]++;
            pos = w & 3;
        }
    }
};

// Block-granular view of the same streams, used by the sampler kernel: every random decision
// point consumes one whole Philox block (four words) and the only state is the block counter.
//   cell choice        block: (w0,w1) -> 53-bit uniform
//   |p| proposal       block: (w0,w1) -> r (53 bit), w2 -> inner accept (32 bit), w3 -> phi (32 bit)
//   direction/accept   block: w0 -> cos(theta) (32 bit), w1 -> accept (32 bit)   [only if the inner
//                      accept passed]
//   rapidity           block: w0 -> y (32 bit)                                   [boost-invariant]
struct BlockStream {
    uint32_t block;     // next block of the stream
    uint32_t draw;      // hadron index inside (event, species)
    uint32_t event;
    ISS_HD void init(uint32_t event_, uint32_t draw_) {
        block = 0;
        draw = draw_;
        event = event_;
    }
};

#if defined(__CUDACC__)
// Out of line on purpose (one copy of the ten unrolled rounds keeps the sampler's hot loop small);
// the four words come back by value, i.e. in registers, not through the local-memory stack.
static __device__ __noinline__ uint4 philox_block4(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                   uint32_t k0, uint32_t k1) {
    const uint32_t ctr[4] = {c0, c1, c2, c3};
    const uint32_t key[2] = {k0, k1};
    uint32_t out[4];
    philox4x32_10(ctr, key, out);
    return make_uint4(out[0], out[1], out[2], out[3]);
}
static __device__ __forceinline__ void philox_block(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                    uint32_t k0, uint32_t k1, uint32_t &w0,
                                                    uint32_t &w1, uint32_t &w2, uint32_t &w3) {
    const uint4 w = philox_block4(c0, c1, c2, c3, k0, k1);
    w0 = w.x; w1 = w.y; w2 = w.z; w3 = w.w;
}
#else
inline void philox_block(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                         uint32_t &w0, uint32_t &w1, uint32_t &w2, uint32_t &w3) {
    const uint32_t ctr[4] = {c0, c1, c2, c3};
    const uint32_t key[2] = {k0, k1};
    uint32_t out[4];
    philox4x32_10(ctr, key, out);
    w0 = out[0]; w1 = out[1]; w2 = out[2]; w3 = out[3];
}
#endif

// Exact Poisson draw by inversion from the mode ("chop-down" outward from
// m = floor(lambda)).  pmode = Poisson pmf at m, computed once per species on the
// host.  Only +,-,*,/ on doubles: bit-identical on CPU and GPU.  Expected number
// of steps ~ 1.6 sqrt(lambda).  lambda < 1e-15 -> 0 as in FSSW.cpp:294-295.
ISS_HD int64_t poisson_from_mode(double lambda, double pmode, double u) {
    if (lambda < 1e-15) return 0;
    int64_t m = static_cast<int64_t>(lambda);
    u = ISS_SUB(u, pmode);
    if (u < 0.0) return m;
    int64_t k_hi = m, k_lo = m;
    double p_hi = pmode, p_lo = pmode;
    for (;;) {
        k_hi++;
        p_hi = ISS_MUL(p_hi, ISS_DIV(lambda, static_cast<double>(k_hi)));
        u = ISS_SUB(u, p_hi);
        if (u < 0.0) return k_hi;
        if (k_lo > 0) {
            p_lo = ISS_MUL(p_lo, ISS_DIV(static_cast<double>(k_lo), lambda));
            k_lo--;
            u = ISS_SUB(u, p_lo);
            if (u < 0.0) return k_lo;
        }
        // numerical tail: both arms exhausted (sum of pmf < 1 by rounding)
        if (p_hi < 1e-300 && (k_lo == 0 || p_lo < 1e-300)) return m;
    }
}

// dN_dy_sampling_model == 1 (FSSW.cpp:269-274): floor + Bernoulli(fraction).
ISS_HD int64_t floor_plus_bernoulli(double dN, double u) {
    int64_t n = static_cast<int64_t>(dN);
    double frac = ISS_SUB(dN, static_cast<double>(n));
    if (u < frac) n++;
    return n;
}

// dN_dy_sampling_model 10 / 20 (FSSW.cpp:275-292): gsl_ran_negative_binomial(p, k) with
// p = 1/(1 + para1), i.e. X ~ Poisson(Y), Y ~ Gamma(shape k, scale (1-p)/p = para1) (GSL's
// definition; GSL itself is not part of the reference tree).  Gamma by Marsaglia & Tsang (2000),
// with the U^(1/k) boost for k < 1; normals by Box-Muller; the Poisson draw is the inversion
// from the mode used everywhere else.  Uses libm transcendentals, so host and device agree
// statistically, not bit for bit.
template <typename RNG>
ISS_HD double gamma_draw(RNG &rng, double shape) {
    double boost = 1.0;
    if (shape < 1.0) {
        boost = pow(rng.next(), 1.0/shape);
        shape += 1.0;
    }
    const double d = shape - 1.0/3.0;
    const double c = 1.0/sqrt(9.0*d);
    for (int it = 0; it < 1000; it++) {
        const double u1 = rng.next(), u2 = rng.next();
        const double x = sqrt(-2.0*log(u1 > 0.0 ? u1 : 1e-300))*cos(6.283185307179586*u2);
        const double t = 1.0 + c*x;
        if (t <= 0.0) continue;
        const double v = t*t*t;
        const double u = rng.next();
        if (log(u > 0.0 ? u : 1e-300) < 0.5*x*x + d - d*v + d*log(v)) return boost*d*v;
    }
    return boost*d;
}

template <typename RNG>
ISS_HD int64_t negative_binomial_draw(RNG &rng, double k, double scale) {
    const double y = scale*gamma_draw(rng, k);
    if (y < 1e-15) return 0;
    const double m = floor(y);
    const double pmode = exp(m*log(y) - y - lgamma(m + 1.0));
    return poisson_from_mode(y, pmode, rng.next());
}

// FSSW::determine_number_to_sample (FSSW.cpp:250-309) for every model
template <typename RNG>
ISS_HD int64_t number_to_sample(RNG &rng, int model, double para1, double dN, double pmode) {
    if (model == 1) return floor_plus_bernoulli(dN, rng.next());
    if (model == 10 || model == 20) {
        const int64_t dN_int = static_cast<int64_t>(dN);
        const double k = (model == 10) ? para1*(dN - static_cast<double>(dN_int)) : para1*dN;
        if (k < 1e-15) return dN_int;
        const int64_t x = negative_binomial_draw(rng, k, para1);
        return (model == 10) ? dN_int + x : x;
    }
    return poisson_from_mode(dN, pmode, rng.next());
}

}  // namespace iss
#endif  // ISS_RNG_H_
