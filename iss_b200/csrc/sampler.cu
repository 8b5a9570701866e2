// sampler.cu -- multiplicities, offsets and the persistent momentum sampler.
//
// Replaces the event/particle loops of
// FSSW::sample_using_dN_dxtdy_4all_particles_conventional (FSSW.cpp:873-1071):
//   K4  multiplicity_kernel : N[ev][s] ~ Poisson(dN_s) per (species, event)
//                             (FSSW::determine_number_to_sample, FSSW.cpp:250-309), Philox
//                             stream (event, species), exact inversion from the mode.
//       + integer exclusive scans (scan.cu) -> output offsets (event-major, species-major
//       inside an event exactly like Hadron_list) and work offsets (species-major).
//   K5  sampler_kernel      : persistent, warp-compacted accept/reject loop.  Every lane owns
//                             one hadron; a warp refills its idle lanes in batches from a
//                             dynamically grabbed chunk of the species-major work list, so that
//                             the proposal loop runs with >= 50 % of the lanes busy regardless of
//                             the acceptance rate.  Per hadron:
//                               cell pick   RandomVariable1DArray::rand + binarySearch
//                                           (RandomVariable1DArray.cpp:63-67, arsenal.cpp:644-678)
//                               |p| draw    MomentumSamplerShell/Base (MomentumSamplerShell.cpp:24-48,
//                                           MomentumSamplerBase.cpp:20-93)
//                               accept      FSSW::sample_momemtum_from_a_fluid_cell (FSSW.cpp:1852-1966)
//                               boost+emit  boost_vector_back_to_lab_frame + add_one_sampled_particle
//                                           (FSSW.cpp:2001-2010, 1969-1996), fused.
// Random numbers: Philox4x32-10 keyed by the seed, counter (block, draw, event, species):
// the hadron list is a pure function of (seed, event, species, draw) and therefore
// identical for any number of GPUs, launch geometry or lane assignment.
#include "coefficients.cuh"

namespace iss {

// ---------------------------------------------------------------------------------
// momentum-sampler tables (Boson/FermionMomentumSampler.cpp)
// ---------------------------------------------------------------------------------
template <bool FERMION>
__host__ __device__ inline void cdf_012(double Et, double m0, int trunc, double &c0, double &c1,
                                        double &c2) {
    // series in n = 0..trunc-1 with e^{-m0 n} and e^{(m0-Et)(n+1)}
    c1 = 0.;
    c2 = 0.;
    double s0 = 0.;
    double sign = 1.;
    for (int n = 0; n < trunc; n++) {
        const int n1 = n + 1;
        const double a = exp(-m0*n);
        const double b = exp((m0 - Et)*n1);
        if (FERMION) {
            // integer division sign/n1 of the reference (FermionMomentumSampler.cpp:43-51):
            // only n = 0 contributes
            if (n == 0) s0 += a*(1. - b);
        } else {
            s0 += (1./n1)*a*(1. - b);
        }
        c1 += ((FERMION ? sign : 1.)/(static_cast<double>(n1)*n1)*a
               *((m0*n1 + 1) - b*(Et*n1 + 1)));
        // CDF_2 carries no alternating sign for fermions either
        // (FermionMomentumSampler.cpp:80-93)
        c2 += (1./(static_cast<double>(n1)*n1*n1)*a
               *((m0*n1*(m0*n1 + 2) + 2) - b*(Et*n1*(Et*n1 + 2) + 2)));
        sign = -sign;
    }
    if (trunc > 5) {
        if (FERMION) {
            c0 = -exp(m0)*log((1. + exp(-Et))/(1. + exp(-m0)));
        } else {
            c0 = exp(m0)*log((1. - exp(-Et))/(1. - exp(-m0)));
        }
    } else {
        c0 = s0;
    }
}

// The same three series at Etilde = a for the per-(cell, species) set-up
// (MomentumSamplerBase::update_cache), with the table constants exp(-m0 n) precomputed and
// exp((m0 - a)(n+1)) formed as powers of exp(m0 - a): 2 exp + 1 log instead of 22 exp + 1 log.
// Differs from the literal series only by FP64 rounding (~1e-16 relative).
template <bool FERMION, bool POLY>
__device__ __forceinline__ void cdf_012_lane(const MomentumTable &mt, double Et, double &c0,
                                             double &c1, double &c2) {
    const double m0 = mt.m0;
    const double e1 = exp(m0 - Et);
    if (POLY && mt.trunc == 10) {
        // the ten terms as polynomials in e1 = exp(m0 - Et) (Horner, 3 + 2 of them):
        //   CDF_2 = k2 - (Et^2 P1(e1) + 2 Et P2(e1) + 2 P3(e1)),  P_j(x) = sum_n a_n/(n+1)^j x^(n+1)
        //   CDF_1 = k1 - (Et P1 + P2) for bosons; the alternating sign of the fermion series is the
        //   same polynomial at -e1: sum_n (-1)^n q_n x^(n+1) = -P(-x)
        // Same terms as the loop below, another order of the additions (~5e-16 of the totals).
        double p1 = mt.q1[9], p2 = mt.q2[9], p3 = mt.q3[9];
#pragma unroll
        for (int n = 8; n >= 0; n--) {
            p1 = __fma_rn(p1, e1, mt.q1[n]);
            p2 = __fma_rn(p2, e1, mt.q2[n]);
            p3 = __fma_rn(p3, e1, mt.q3[n]);
        }
        p1 *= e1;
        p2 *= e1;
        p3 *= e1;
        c2 = mt.k2 - (Et*Et*p1 + 2.*Et*p2 + 2.*p3);
        if (FERMION) {
            const double x = -e1;
            double s1 = mt.q1[9], s2 = mt.q2[9];
#pragma unroll
            for (int n = 8; n >= 0; n--) {
                s1 = __fma_rn(s1, x, mt.q1[n]);
                s2 = __fma_rn(s2, x, mt.q2[n]);
            }
            c1 = mt.k1 - (Et*(s1*e1) + s2*e1);      // -P(-e1) = (P(x)/x at x = -e1) e1
            c0 = -mt.exp_m0*log((1. + e1*mt.a[1])*mt.inv_denom0);
        } else {
            c1 = mt.k1 - (Et*p1 + p2);
            c0 = mt.exp_m0*log((1. - e1*mt.a[1])*mt.inv_denom0);
        }
        return;
    }
    double b = 1.;
    double s0 = 0.;
    c1 = 0.;
    c2 = 0.;
    double sign = 1.;
#pragma unroll 1
    for (int n = 0; n < mt.trunc; n++) {
        const double n1 = n + 1;
        b *= e1;
        const double a = mt.a[n];
        const double inv = mt.inv_n1[n];        // 1/(n+1)
        if (FERMION) {
            if (n == 0) s0 += a*(1. - b);
        } else {
            s0 += inv*a*(1. - b);
        }
        c1 += ((FERMION ? sign : 1.)*inv*inv*a*((m0*n1 + 1) - b*(Et*n1 + 1)));
        c2 += (inv*inv*inv*a*((m0*n1*(m0*n1 + 2) + 2) - b*(Et*n1*(Et*n1 + 2) + 2)));
        sign = -sign;
    }
    if (mt.trunc > 5) {
        // exp(-Et) = exp(m0 - Et) exp(-m0); log(x/denom) with 1/denom precomputed
        const double emE = e1*mt.a[1];
        if (FERMION) {
            c0 = -mt.exp_m0*log((1. + emE)*mt.inv_denom0);
        } else {
            c0 = mt.exp_m0*log((1. - emE)*mt.inv_denom0);
        }
    } else {
        c0 = s0;
    }
}

__global__ void build_momentum_table_kernel(double *tab, int n, double e0, double de, double m0,
                                            int trunc, int fermion) {
    const int i = blockIdx.x*blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double Et = __fma_rn(static_cast<double>(i), de, e0);     // = e0 + i*de as compiled so far
    double c0, c1, c2;
    if (fermion) {
        cdf_012<true>(Et, m0, trunc, c0, c1, c2);
    } else {
        cdf_012<false>(Et, m0, trunc, c0, c1, c2);
    }
    tab[i*4 + 0] = Et;
    tab[i*4 + 1] = c0;
    tab[i*4 + 2] = c1;
    tab[i*4 + 3] = c2;
}

static int ensure_momentum_tables(iss_handle *h) {
    // MomentumSamplerShell ctor (MomentumSamplerShell.cpp:6-21)
    const double m0_b[3] = {0.05, 30., 50.};
    const double m0_f[3] = {0., 30., 50.};
    const int trunc[3] = {10, 2, 1};
    for (int r = 0; r < 6; r++) {
        if (h->d_momtab[r]) continue;
        const bool fermion = r >= 3;
        const double m0 = fermion ? m0_f[r - 3] : m0_b[r];
        // same double arithmetic as the reference constructors
        double E_min, E_max, dE;
        if (fermion) {
            E_min = m0;
            E_max = E_min + 40.;
            dE = 0.02;
        } else {
            E_min = m0 + 0.05;
            E_max = E_min + 50.;
            dE = 0.05;
        }
        const int npoints = (E_max - E_min)/dE + 1;
        ISS_CUDA_TRY(h, cudaMalloc(&h->d_momtab[r], sizeof(double)*4*npoints));
        build_momentum_table_kernel<<<(npoints + 127)/128, 128, 0, h->stream>>>(
            h->d_momtab[r], npoints, E_min, dE, m0, trunc[r % 3], fermion ? 1 : 0); ISS_LAUNCHED(h);
        ISS_CUDA_TRY(h, cudaGetLastError());
        MomentumTable &t = h->momtab[r];
        t.data = h->d_momtab[r];
        t.n = npoints;
        t.trunc = trunc[r % 3];
        t.m0 = m0;
        t.e0 = E_min;
        t.de = (E_min + dE) - E_min;    // Etilde_[1] - Etilde_[0]
        t.de_build = dE;
        t.generated = 1;
        t.exp_m0 = exp(m0);
        t.denom0 = fermion ? (1. + exp(-m0)) : (1. - exp(-m0));
        t.inv_denom0 = 1.0/t.denom0;
        t.inv_de = 1.0/t.de;
        momentum_series_constants(t, fermion);
    }
    return ISS_OK;
}

// ---------------------------------------------------------------------------------
// K4: multiplicities
// ---------------------------------------------------------------------------------
__global__ void multiplicity_kernel(const double *__restrict__ lambda,
                                    const double *__restrict__ pmode,
                                    const DeviceSpecies *__restrict__ species, int ns, int64_t nev,
                                    int64_t ev_begin, uint64_t seed, int model, double para1, int lcc,
                                    int64_t *__restrict__ mult,       // [nev][ns] draws
                                    int64_t *__restrict__ out_count,  // [nev*ns] hadrons written
                                    int64_t *__restrict__ work_count  // [ns*nev] draws
) {
    const int64_t i = static_cast<int64_t>(blockIdx.x)*blockDim.x + threadIdx.x;
    if (i >= nev*ns) return;
    const int64_t ev = i/ns;
    const int s = static_cast<int>(i - ev*ns);
    Stream rng;
    rng.init(seed, STREAM_MULT, s, static_cast<uint32_t>(ev_begin + ev), 0);
    int64_t n = number_to_sample(rng, model, para1, lambda[s], pmode[s]);
    int64_t nout = n;
    if (lcc == 1) {
        // FSSW.cpp:931-938, 1035-1048: negative species skipped, positive ones paired
        const int q = species[s].charge;
        if (q < 0) {
            n = 0;
            nout = 0;
        } else if (q > 0) {
            nout = 2*n;
        }
    }
    mult[i] = n;
    out_count[i] = nout;
    work_count[static_cast<int64_t>(s)*nev + ev] = n;
}

// per-event offsets = off_out[ev*ns]
__global__ void event_offset_kernel(const int64_t *__restrict__ off_out, int ns, int64_t nev,
                                    int64_t *__restrict__ event_off,
                                    int64_t *__restrict__ event_off_host /* mapped pinned */) {
    const int64_t ev = static_cast<int64_t>(blockIdx.x)*blockDim.x + threadIdx.x;
    if (ev <= nev) {
        const int64_t v = off_out[ev*ns];
        event_off[ev] = v;
        event_off_host[ev] = v;
    }
}

// ---------------------------------------------------------------------------------
// K5: sampler = cell pick -> cell-sorted task list -> persistent proposal kernel
// ---------------------------------------------------------------------------------
// One hadron to sample: 32 bytes (one DRAM sector), written by setup_kernel at its position in
// the CELL-SORTED task list and streamed into shared memory by propose_kernel (cp.async.bulk).
// The output slot follows from (event, species, draw), so results do not depend on the order.
struct __align__(16) Task32 {
    double m_term;          // CDF(a) term of MomentumSamplerBase::Sample_a_momentum
    double cdf_max;
    uint32_t cell;          // (local) cell index
    uint32_t event;         // global event index
    uint32_t draw;          // index of the hadron inside (event, species)
    uint16_t s;             // species (sampling order)
    uint16_t tab_idx;       // momentum table (bits 0..2) | idx_min << 3; TASK_RANGE_ERROR: table range error
};
static_assert(sizeof(Task32) == 32, "Task32 must stay 32 bytes");
constexpr uint16_t TASK_RANGE_ERROR = 0xFFFFu;

// Everything the proposal kernel needs from a cell, in final form: the fields of the try, the
// emit fields and every quotient of the accept test that does not depend on the proposed momentum
// (the reference recomputes them per sample, FSSW.cpp:1861-1913).  Built once per yields
// computation by build_cellrec_kernel; 160 bytes = 10 chunks of 16 bytes, chunks 0..8 are copied
// verbatim into the lane's shared-memory slot, chunk 9 is consumed at task hand-over.
struct __align__(16) CellRec {
    float4 da;              // [0] dsigma_mu (LRF)
    float4 pa;              // [1] pixx, pixy, pixz, piyy
    float4 pb;              // [2] piyz, qx, qy, qz
    float4 u4;              // [3] ut, ux, uy, uz
    float4 pos;             // [4] tau, x, y, eta
    float2 tz;              // [5] t, z of the cell
    double inv_T;           //     1/max(T, 1e-16)
    double inv_dsig;        // [6] 1/(|dsigma_0| + |dsigma_vec|)
    double shear;           //     shear delta-f prefactor (FSSW.cpp:1898-1913)
    double cb;              // [7] c0 * Pi of the CE bulk delta f (kinds 1, 21)
    double c1;
    double inv_kappa;       // [8] 1/kappa_B
    double prefq;           //     n_B/(e + P) (float division, FSSW.cpp:1866)
    float4 th;              // [9] T, muB, muS, muQ
};
static_assert(sizeof(CellRec) == 160, "CellRec must stay 160 bytes");
constexpr int CELLREC_SLOT_CHUNKS = 9;

// what the proposal kernel reads per species (shared memory, 32 B)
struct __align__(16) PropSpecies {
    double mass, mass2;
    int32_t pid;
    int16_t baryon, strange, charge, sign;
    int32_t pad;
};
static_assert(sizeof(PropSpecies) == 32, "PropSpecies must stay 32 bytes");

struct SamplerArgs {
    const float *cells;             // [ncell][CELL_STRIDE]
    const float4 *thermo;           // [ncell] {T, muB, muS, muQ}
    const double *cellcoef;         // [ncell][COEF_STRIDE]
    int64_t ncell, ncell_pad;
    const double *cdf;              // [ns][ncell_pad] global inclusive prefix of the yields
    const double *cdflev;           // [ns][lev_stride] 16-ary search levels over cdf
    int nlev;
    int64_t lev_off[8], lev_stride;
    const double *total;            // [ns]
    const int2 *guide;              // [ns][guide_M + 1] {G[k], G[k+1]} over level 1, or null
    int64_t guide_M;
    const DeviceSpecies *species;
    int ns;
    int64_t nev, ev_begin;
    const int64_t *off_work;        // [ns*nev+1]
    const int64_t *off_out;         // [nev*ns+1]
    int64_t nwork;
    MomentumTable mt[6];
    ModeFlags mode;
    int hydro_mode, lcc;
    double y_LB, y_RB;
    uint64_t seed;
    const CellRec *cellrec;         // [ncell] per-cell record of the proposal kernel
    Task32 *tasks_unsorted;         // [nwork] tasks in work order (setup_kernel -> partition_kernel)
    unsigned long long *bucket_cnt; // [nbucket][nseg] (+1) tasks per (cell block, segment of PART_TILE
                                    // work items) -> exclusive scan: write offset of every run
    int bucket_shift, nbucket;      // block = cell >> bucket_shift
    int64_t nseg;                   // segments of the batch
    Task32 *tasks;                  // [nstage*RING_TASKS] cell-sorted task list
    uint32_t *task_slot;            // surface-chunk mode only: output slot of every sorted task
    unsigned long long *giveup_info;    // [2] (cell, species) of a hadron the sampler gave up on
    int mt_smem;                    // regime-0 tables were generated on the device (uniform abscissa)
    const int2 *hints;              // [ceil(nwork/SETUP_THREADS)] (species, event) of item b*SETUP_THREADS
    iss_hadron *out;
    unsigned long long *counters;   // [0] task cursor (legacy), [1] tries, [2] redraws, [3] range errors,
                                    // [4],[5] decay errors, [6] hadrons given up, [7] foreign re-draws
    int32_t *trace_cell;            // optional [n_out]
    int32_t *trace_tries;           // optional [n_out]
    // surface-chunk mode (iss_cuda_set_surface_chunk): cells/cdf/cdflev above are those of the
    // local cells; levels >= 3 of the search tree over the whole surface are in cdflev_g
    int chunk;
    int g_nlev;
    const double *cdflev_g;         // [ns][g_lev_stride]
    int64_t g_lev_off[8], g_lev_stride;
    int64_t blk_begin, blk_end;     // 4096-cell blocks of the whole surface this rank owns
    int64_t cell_begin, g_ncell;
    const uint4 *ident;             // [nwork] (species, event - ev_begin, draw, output slot) of the rank's hadrons
};

constexpr int SETUP_THREADS = 256;
constexpr int PART_THREADS = 256;       // partition_kernel: tasks of a segment per thread
constexpr int PART_ITEMS = 8;
constexpr int PART_TILE = PART_THREADS*PART_ITEMS;      // work items per segment
constexpr int MAX_BUCKETS = 64;         // cell blocks of the partition (partition_kernel scans them with one warp)
static_assert(MAX_BUCKETS == 64, "partition_kernel scans two counters per lane of one warp");
constexpr int SAMPLER_THREADS = 768;     // one CTA of 24 warps per SM (80 registers per thread)
constexpr int MAX_IMPATIENCE = 5000;    // FSSW.cpp:1872
constexpr int MAX_TRIES_PER_HADRON = 2000000;   // safety valve, see propose_kernel

// per-(cell, species) constants of the |p| sampler (MomentumSamplerBase::Sample_a_momentum)
struct MomSetup {
    double T, mu, mu_tilde, w0, m_term, cdf_max, a_min;
    int tab, idx_min;
};

// F(i) = CDF_2[i] + w1 CDF_1[i] + w0 CDF_0[i] - m_term of one table point; SMEM: the table is the
// CTA's shared-memory copy (regime-0 tables), otherwise global memory through the read-only path
template <bool SMEM>
__device__ __forceinline__ double table_F(const double *__restrict__ tb, int i, double w1,
                                          double w0, double m_term) {
    double2 a, b;
    if (SMEM) {
        a = *reinterpret_cast<const double2 *>(tb + 4*i);
        b = *reinterpret_cast<const double2 *>(tb + 4*i + 2);
    } else {
        a = __ldg(reinterpret_cast<const double2 *>(tb + 4*i));       // Et, C0
        b = __ldg(reinterpret_cast<const double2 *>(tb + 4*i + 2));   // C1, C2
    }
    return b.y + w1*b.x + w0*a.y - m_term;
}

// mu/T.  0/T = 0 exactly, but a zero numerator sends the FP64 division down its slow path (a quarter
// of the hadrons: every species with B = S = Q = 0), so a harmless numerator is divided instead; the
// empty asm keeps the compiler from folding the two selects back into one division.
__device__ __forceinline__ double mu_over_T(double mu, double T) {
    double num = (mu == 0.) ? 1. : mu;
    asm volatile("" : "+d"(num));
    const double q = num/T;
    return (mu == 0.) ? mu : q;
}

// the cheap part: everything but the two series values, which travel in the Task
__device__ __forceinline__ void momentum_restore(double mass, double T_in, double mu, double m_term,
                                                 double cdf_max, int tab, int idx_min, MomSetup &M) {
    const double T = fmax(1e-16, T_in);
    const double m_tilde = mass/T;
    // 0/T = 0 exactly, but a zero numerator sends the FP64 division down its slow path (a quarter
    // of the hadrons: every species with B = S = Q = 0): divide a harmless numerator instead
    const double mu_tilde = mu_over_T(mu, T);
    M.T = T;
    M.mu = mu;
    M.mu_tilde = mu_tilde;
    M.w0 = mu_tilde*mu_tilde - m_tilde*m_tilde/2.;
    M.m_term = m_term;
    M.cdf_max = cdf_max;
    M.a_min = m_tilde - mu_tilde;
    M.tab = tab;
    M.idx_min = idx_min;
}

// full set-up for (species, cell); returns false if (m - mu)/T is outside the table
// (reference: exit(1), MomentumSamplerBase.cpp:35-43)
// POLY: the polynomial form of the ten-term series (the set-up kernel, one call per hadron); the
// rare paths inside the proposal kernel keep the rolled loop, which costs that kernel no registers
template <bool POLY = false>
__device__ __forceinline__ bool momentum_setup(const MomentumTable *__restrict__ mts, double mass,
                                               int sign, double T_in, double mu, MomSetup &M) {
    const double T = fmax(1e-16, T_in);
    const double m_tilde = mass/T;
    const double mu_tilde = mu_over_T(mu, T);
    const double a = m_tilde - mu_tilde;
    const double m0tilde = m_tilde - mu_tilde;
    const int regime = (m0tilde < 30.) ? 0 : (m0tilde < 50. ? 1 : 2);
    const bool fermion = (sign != -1);          // sign 0 -> fermion tables
    const int tab = (fermion ? 3 : 0) + regime;
    const MomentumTable &mt = mts[tab];
    double c0, c1, c2;
    if (fermion) {
        cdf_012_lane<true, POLY>(mt, a, c0, c1, c2);
    } else {
        cdf_012_lane<false, POLY>(mt, a, c0, c1, c2);
    }
    const double w1 = 2.*mu_tilde;
    const double w0 = mu_tilde*mu_tilde - m_tilde*m_tilde/2.;
    const double m_term = c2 + w1*c1 + w0*c0;
    const int idx_max = mt.n - 1;
    M.T = T;
    M.mu = mu;
    M.mu_tilde = mu_tilde;
    M.w0 = w0;
    M.m_term = m_term;
    M.cdf_max = table_F<false>(mt.data, idx_max, w1, w0, m_term);
    M.a_min = a;
    M.tab = tab;
    // int((a - Etilde_0)/dEtilde) (MomentumSamplerBase.cpp:33-34) with the reciprocal spacing; when
    // the product lands within rounding of an integer the exact quotient decides
    {
        const double t = (a - mt.e0)*mt.inv_de;
        int idx = static_cast<int>(t);
        if (fabs(t - rint(t)) < 1e-9) idx = static_cast<int>((a - mt.e0)/mt.de);
        M.idx_min = idx;
    }
    return !(M.idx_min < 0 || M.idx_min >= idx_max);
}

// MomentumSamplerBase::inverse_CDF (MomentumSamplerBase.cpp:61-93)
template <bool SMEM>
__device__ __forceinline__ double inverse_cdf(const double *__restrict__ tb, int n, const MomSetup &M,
                                              double r) {
    const double w1 = 2.*M.mu_tilde;
    int lo = M.idx_min;
    int hi = n - 1;
    double r_min = table_F<SMEM>(tb, lo, w1, M.w0, M.m_term);
    double r_max = M.cdf_max;
    while (hi - lo > 1) {
        const int mid = (hi + lo)/2;
        const double r_mid = table_F<SMEM>(tb, mid, w1, M.w0, M.m_term);
        if (r < r_mid) {
            hi = mid;
            r_max = r_mid;
        } else {
            lo = mid;
            r_min = r_mid;
        }
    }
    double E0 = SMEM ? tb[4*lo] : __ldg(tb + 4*lo);
    if (E0 < M.a_min) {
        E0 = M.a_min;
        r_min = 0.;
    }
    const double Ehi = SMEM ? tb[4*hi] : __ldg(tb + 4*hi);
    return E0 + (Ehi - E0)/fmax(1e-16, (r_max - r_min))*(r - r_min);
}

// number of entries < v among the 16 sorted doubles of one 128-byte node whose last entry is
// known to be >= v: four dependent loads inside one cache line
__device__ __forceinline__ int count_below_16(const double *__restrict__ node, double v) {
    int c = 0;
    c += (__ldg(&node[c + 7]) < v) ? 8 : 0;
    c += (__ldg(&node[c + 3]) < v) ? 4 : 0;
    c += (__ldg(&node[c + 1]) < v) ? 2 : 0;
    c += (__ldg(&node[c]) < v) ? 1 : 0;
    return c;
}

// RandomVariable1DArray::rand (RandomVariable1DArray.cpp:63-67): v = (sum - 1e-15) u, cell =
// largest i with CDF[i] < v where CDF is the exclusive prefix, i.e. the number of inclusive-prefix
// entries below v.  The reference bisects the whole array (arsenal.cpp:644-678); here the same
// count is obtained by descending a 16-ary tree of sampled prefix values: one 128-byte node per
// level, five levels for 10^6 cells.
__device__ __forceinline__ int64_t pick_cell(const SamplerArgs &A, int s, double u) {
    const double total = __ldg(&A.total[s]);
    const double v = (total - 1e-15)*u;
    const double *__restrict__ lev = A.cdflev + static_cast<int64_t>(s)*A.lev_stride;
    int64_t q = 0;
    if (A.chunk) {
        // the same descent through the same numbers as on one GPU: levels >= 3 over the whole
        // surface pick the 4096-cell block; a foreign block ends the search (-1), an owned one is
        // continued in the local levels 2, 1 and the local prefix
        const double *__restrict__ levg = A.cdflev_g + static_cast<int64_t>(s)*A.g_lev_stride;
        for (int k = A.g_nlev; k >= 3; k--) q = 16*q + count_below_16(levg + A.g_lev_off[k] + 16*q, v);
        if (q < A.blk_begin || q >= A.blk_end) return -1;
        q -= A.blk_begin;
        q = 16*q + count_below_16(lev + A.lev_off[2] + 16*q, v);
        q = 16*q + count_below_16(lev + A.lev_off[1] + 16*q, v);
        const double *__restrict__ Pl = A.cdf + static_cast<int64_t>(s)*A.ncell_pad;
        int64_t cl = 16*q + count_below_16(Pl + 16*q, v);
        if (cl + A.cell_begin >= A.g_ncell) cl = A.g_ncell - 1 - A.cell_begin;
        return cl;
    }
    if (A.guide) {
        // guide table (yields.cu, guide_kernel): u M is exact (M a power of two), the number of
        // level-1 entries below v lies in [G[k], G[k+1]]
        const int64_t k = static_cast<int64_t>(u*static_cast<double>(A.guide_M));
        const int2 g = __ldg(&A.guide[static_cast<int64_t>(s)*(A.guide_M + 1) + k]);
        const double *__restrict__ L1 = lev + A.lev_off[1];
        int lo = g.x, hi = g.y;
        while (hi - lo > 4) {               // wide brackets (flat stretches of the prefix): bisect
            const int mid = (lo + hi) >> 1;
            if (__ldg(&L1[mid]) < v) lo = mid + 1; else hi = mid;
        }
        while (lo < hi && __ldg(&L1[lo]) < v) lo++;
        q = lo;
    } else {
        for (int k = A.nlev; k >= 1; k--) q = 16*q + count_below_16(lev + A.lev_off[k] + 16*q, v);
    }
    const double *__restrict__ P = A.cdf + static_cast<int64_t>(s)*A.ncell_pad;
    int64_t cell = 16*q + count_below_16(P + 16*q, v);
    if (cell >= A.ncell) cell = A.ncell - 1;
    return cell;
}

// chemical potential of a species in a cell: float arithmetic and min(m, mu) as FSSW.cpp:1861-1863
__device__ __forceinline__ double species_mu(const DeviceSpecies &p, int qsign, const float4 th,
                                             double mass) {
    const float muf = __fadd_rn(
        __fadd_rn(__fmul_rn(static_cast<float>(qsign*p.baryon), th.x),
                  __fmul_rn(static_cast<float>(qsign*p.strange), th.y)),
        __fmul_rn(static_cast<float>(qsign*p.charge), th.z));
    return fmin(mass, static_cast<double>(muf));
}

__device__ __forceinline__ double species_mu_bsq(int B, int S, int Q, float muB, float muS, float muQ,
                                                 double mass) {
    const float muf = __fadd_rn(__fadd_rn(__fmul_rn(static_cast<float>(B), muB),
                                          __fmul_rn(static_cast<float>(S), muS)),
                                __fmul_rn(static_cast<float>(Q), muQ));
    return fmin(mass, static_cast<double>(muf));
}

__device__ __forceinline__ uint32_t sample_stream_word3(int s) {
    return (static_cast<uint32_t>(STREAM_SAMPLE) << 24) | static_cast<uint32_t>(s);
}

// (species, event) of every SETUP_THREADS-th work item by full binary searches; setup_kernel's
// threads start from these hints
__global__ void work_hint_kernel(const int64_t *__restrict__ off_work, int ns, int64_t nev,
                                 int64_t nwork, int2 *__restrict__ hints, int64_t nhint) {
    const int64_t b = static_cast<int64_t>(blockIdx.x)*blockDim.x + threadIdx.x;
    if (b >= nhint) return;
    const int64_t w = min(b*SETUP_THREADS, nwork - 1);
    int slo = 0, shi = ns;
    while (shi - slo > 1) {
        const int mid = (slo + shi) >> 1;
        if (__ldg(&off_work[static_cast<int64_t>(mid)*nev]) <= w) slo = mid; else shi = mid;
    }
    const int64_t *__restrict__ ow = off_work + static_cast<int64_t>(slo)*nev;
    int64_t elo = 0, ehi = nev;
    while (ehi - elo > 1) {
        const int64_t mid = (elo + ehi) >> 1;
        if (__ldg(&ow[mid]) <= w) elo = mid; else ehi = mid;
    }
    hints[b] = make_int2(slo, static_cast<int>(elo));
}

// (species, event, draw) of work item w: start from the hint stored for the first item of its group
// of SETUP_THREADS consecutive items and move forward (work items are species-major, event-minor,
// so the answer is at or after the hint).  sp_off[s] = off_work[s*nev] (shared memory).
__device__ __forceinline__ void work_identity(const SamplerArgs &A, const int64_t *sp_off, int64_t w,
                                              int &s_out, int64_t &ev_out, int64_t &k_out) {
    const int2 hint = __ldg(&A.hints[w/SETUP_THREADS]);
    int s = hint.x;
    while (sp_off[s + 1] <= w) s++;
    const int64_t *__restrict__ ow = A.off_work + static_cast<int64_t>(s)*A.nev;
    int64_t elo = (s == hint.x) ? hint.y : 0, step = 1;
    while (elo + step < A.nev && __ldg(&ow[elo + step]) <= w) {
        elo += step;
        step <<= 1;
    }
    int64_t ehi = min(elo + step, A.nev);
    while (ehi - elo > 1) {
        const int64_t mid = (elo + ehi) >> 1;
        if (__ldg(&ow[mid]) <= w) elo = mid; else ehi = mid;
    }
    s_out = s;
    ev_out = elo;
    k_out = w - __ldg(&ow[elo]);
}

// Surface-chunk mode: which draws of the batch land in this rank's cells.  pick_cell counts the
// prefix entries below v; the entries of level 3 of the tree are the inclusive prefix at the end
// of every 4096-cell block, and the prefix is monotone, so "the block of the descent lies in
// [blk_begin, blk_end)" is the same statement as P_end[blk_begin - 1] < v <= P_end[blk_end - 1]: two
// comparisons with numbers every rank holds (chunk_level3_kernel), no descent, no per-hadron
// arrays.  One warp per (species, event) pair, lanes over its draws (one Philox block each):
// owned draws per pair -> cnt[s*nev + ev], the rank's per-(event, species) output counts ->
// off_out[ev*ns + s] (unscanned), and one ballot word per 32 draws -> mask (word chunk_mask_word
// (j) + i for draws 32 i .. 32 i + 31 of pair j) for chunk_fill_kernel.
__device__ __forceinline__ int64_t chunk_mask_word(const int64_t *__restrict__ off_work, int64_t j) {
    // pair j needs ceil(n_j/32) words; floor(off_work[j]/32) + j never overlaps the next pair's range
    return (__ldg(&off_work[j]) >> 5) + j;
}

__global__ void __launch_bounds__(256)
chunk_own_kernel(const SamplerArgs A, int64_t *__restrict__ cnt, int64_t *__restrict__ off_out,
                 uint32_t *__restrict__ mask) {
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const uint32_t key0 = static_cast<uint32_t>(A.seed), key1 = static_cast<uint32_t>(A.seed >> 32);
    const int64_t npair = static_cast<int64_t>(A.ns)*A.nev;
    const int64_t nwarp = static_cast<int64_t>(gridDim.x)*(blockDim.x >> 5);
    const bool has_lo = A.blk_begin > 0, has_hi = A.blk_end < (int64_t(1) << 61);
    for (int64_t j = static_cast<int64_t>(blockIdx.x)*(blockDim.x >> 5) + (threadIdx.x >> 5); j < npair;
         j += nwarp) {
        const int64_t n = __ldg(&A.off_work[j + 1]) - __ldg(&A.off_work[j]);
        const int s = static_cast<int>(j/A.nev);
        const int64_t ev = j - static_cast<int64_t>(s)*A.nev;
        const int mult = (A.lcc == 1 && __ldg(&A.species[s].charge) > 0) ? 2 : 1;
        int64_t running = 0;
        if (n > 0) {
            const double *__restrict__ l3 = A.cdflev_g + static_cast<int64_t>(s)*A.g_lev_stride + A.g_lev_off[3];
            const double lo = has_lo ? __ldg(&l3[A.blk_begin - 1]) : 0.;
            const double hi = has_hi ? __ldg(&l3[A.blk_end - 1]) : 0.;
            const double scale = __ldg(&A.total[s]) - 1e-15;
            const uint32_t event = static_cast<uint32_t>(A.ev_begin + ev);
            uint32_t *__restrict__ mw = mask + chunk_mask_word(A.off_work, j);
            for (int64_t k0 = 0; k0 < n; k0 += 32) {
                const int64_t k = k0 + lane;
                bool own = false;
                if (k < n) {
                    uint32_t w0, w1, w2, w3;
                    philox_block(0u, static_cast<uint32_t>(k), event, sample_stream_word3(s), key0, key1,
                                 w0, w1, w2, w3);
                    const double v = scale*u53(w0, w1);
                    own = (!has_lo || lo < v) && (!has_hi || !(hi < v));
                }
                const unsigned m = __ballot_sync(full, own);
                if (lane == 0) mw[k0 >> 5] = m;
                running += __popc(m);
            }
        }
        if (lane == 0) {
            cnt[j] = running;
            off_out[ev*A.ns + s] = running*mult;
        }
    }
}

// after the scans of cnt (own_off) and off_out: the identity (species, event, draw, output slot) of
// every owned draw in work order (species-major, event, draw) from the ballot words.  One warp per
// pair, one word per lane, the set bits of a word written one after the other.
__global__ void __launch_bounds__(256)
chunk_fill_kernel(const SamplerArgs A, const int64_t *__restrict__ own_off,
                  const int64_t *__restrict__ off_out, const uint32_t *__restrict__ mask,
                  uint4 *__restrict__ ident) {
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int64_t npair = static_cast<int64_t>(A.ns)*A.nev;
    const int64_t nwarp = static_cast<int64_t>(gridDim.x)*(blockDim.x >> 5);
    for (int64_t j = static_cast<int64_t>(blockIdx.x)*(blockDim.x >> 5) + (threadIdx.x >> 5); j < npair;
         j += nwarp) {
        const int64_t base = __ldg(&own_off[j]);
        if (__ldg(&own_off[j + 1]) == base) continue;       // nothing of this pair is ours
        const int64_t n = __ldg(&A.off_work[j + 1]) - __ldg(&A.off_work[j]);
        const int s = static_cast<int>(j/A.nev);
        const int64_t ev = j - static_cast<int64_t>(s)*A.nev;
        const int mult = (A.lcc == 1 && __ldg(&A.species[s].charge) > 0) ? 2 : 1;
        const int64_t slot0 = __ldg(&off_out[ev*A.ns + s]);
        const uint32_t *__restrict__ mw = mask + chunk_mask_word(A.off_work, j);
        const int64_t nword = (n + 31) >> 5;
        int64_t running = 0;
        for (int64_t i0 = 0; i0 < nword; i0 += 32) {
            const int64_t i = i0 + lane;
            uint32_t m = (i < nword) ? __ldg(&mw[i]) : 0u;
            const int c = __popc(m);
            int incl = c;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int x = __shfl_up_sync(full, incl, d);
                if (lane >= d) incl += x;
            }
            int64_t r = running + (incl - c);
            while (m) {
                const int b = __ffs(m) - 1;
                m &= m - 1u;
                ident[base + r] = make_uint4(static_cast<uint32_t>(s), static_cast<uint32_t>(ev),
                                             static_cast<uint32_t>(32*i + b),
                                             static_cast<uint32_t>(slot0 + r*mult));
                r++;
            }
            running += __shfl_sync(full, incl, 31);
        }
    }
}

// per-cell record of the proposal kernel from the AoS cell record and the delta-f coefficients
__global__ void build_cellrec_kernel(const float *__restrict__ cells, const double *__restrict__ cellcoef,
                                     int64_t ncell, ModeFlags mode, CellRec *__restrict__ out) {
    const int64_t c = static_cast<int64_t>(blockIdx.x)*blockDim.x + threadIdx.x;
    if (c >= ncell) return;
    const float4 *cr = reinterpret_cast<const float4 *>(cells + c*CELL_STRIDE);
    const float4 pos = __ldg(cr + 0), da = __ldg(cr + 1), u4 = __ldg(cr + 2);
    const float4 th0 = __ldg(cr + 3);       // E, T, P, nB
    const float4 th = __ldg(cr + 4);        // muB, muS, muQ, bulkPi
    const float4 tzq = __ldg(cr + 7);       // t, z, spare, spare
    const double2 *cop = reinterpret_cast<const double2 *>(cellcoef + c*COEF_STRIDE);
    const double2 c01 = __ldg(cop);
    const double c2 = __ldg(&cop[1].x), kappa = __ldg(&cop[3].x);
    CellRec r;
    r.da = da;
    r.pa = __ldg(cr + 5);
    r.pb = __ldg(cr + 6);
    r.u4 = u4;
    r.pos = pos;
    r.tz = make_float2(tzq.x, tzq.y);
    const double Tdec = th0.y;
    r.inv_T = 1.0/fmax(1e-16, Tdec);
    // float arithmetic inside the sqrt as in FSSW.cpp:1867-1870
    const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(da.y, da.y), __fmul_rn(da.z, da.z)),
                               __fmul_rn(da.w, da.w));
    r.inv_dsig = 1.0/(fabs(static_cast<double>(da.x)) + sqrt(static_cast<double>(d2)));
    // shear delta f prefactor (FSSW.cpp:1898-1913): CE W/(2 eta_hat p0 T), 22-moment W c0,
    // otherwise W/(2 T^2 (e+P))
    if (mode.neos == 1) r.shear = 1.0/(2.*c2);
    else if (mode.neos == 0) r.shear = c01.x;
    else r.shear = 1.0/(2.0*Tdec*Tdec*(static_cast<double>(__fadd_rn(th0.x, th0.z))));
    // CE bulk delta f (kinds 1, 21): c0 * Pi with Pi in GeV/fm^3 (21) or fm^-4 (1)
    const double bulkPi = (mode.kind == 21) ? static_cast<double>(th.w)
                                            : static_cast<double>(th.w)/HBARC;
    r.cb = c01.x*bulkPi;
    r.c1 = c01.y;
    r.inv_kappa = 1.0/kappa;
    r.prefq = __fdiv_rn(th0.w, __fadd_rn(th0.x, th0.z));   // float division as in FSSW.cpp:1866
    r.th = make_float4(th0.y, th.x, th.y, th.z);
    out[c] = r;
}

// ---- cell-sorted task list -----------------------------------------------------------------------
// The hadron list is a pure function of (seed, event, species, draw), so the ORDER in which the
// hadrons of a batch are sampled is free.  The tasks are partitioned by cell block (a few hundred
// blocks of 2^k consecutive cells: histogram in setup_kernel, exclusive scan, scatter in
// partition_kernel): the tasks in flight at any time then share a few hundred KB of cell
// records, which the proposal kernel finds in L2 instead of gathering 160 random bytes per hadron
// from DRAM.  The order inside a block depends on atomic timing; the results do not (the output
// slot follows from the task).

// K5a: one thread per hadron of the batch, in work order: identity (species, event, draw) from the
// species-major work offsets, cell choice (first block of the hadron's stream), the two series
// values of the |p| sampler (MomentumSamplerBase::update_cache), histogram of the cells.  The FP64
// work of the series runs in the shadow of the dependent loads of the cell search.  The task is
// written at its work index; partition_kernel moves it into the region of its cell block.
__global__ void __launch_bounds__(SETUP_THREADS, 5)
setup_kernel(const SamplerArgs A) {
    extern __shared__ unsigned char smem_raw[];
    DeviceSpecies *sp = reinterpret_cast<DeviceSpecies *>(smem_raw);
    int64_t *sp_off = reinterpret_cast<int64_t *>(sp + A.ns);     // off_work[s*nev], s = 0..ns
    unsigned int *hist = reinterpret_cast<unsigned int *>(sp_off + A.ns + 1);      // [nbucket]
    for (int i = threadIdx.x; i < A.ns; i += blockDim.x) sp[i] = A.species[i];
    for (int i = threadIdx.x; i <= A.ns; i += blockDim.x)
        sp_off[i] = A.off_work[static_cast<int64_t>(i)*A.nev];
    for (int b = threadIdx.x; b < A.nbucket; b += blockDim.x) hist[b] = 0u;
    __syncthreads();
    const uint32_t key0 = static_cast<uint32_t>(A.seed), key1 = static_cast<uint32_t>(A.seed >> 32);
    unsigned long long my_range = 0;
    // a CTA works through segments of PART_TILE consecutive work items (the tiles of
    // partition_kernel) and leaves the histogram of every segment over the cell blocks
    for (int64_t seg = blockIdx.x; seg < A.nseg; seg += gridDim.x) {
      for (int it = 0; it < PART_TILE/SETUP_THREADS; it++) {
        const int64_t j = seg*PART_TILE + it*SETUP_THREADS + threadIdx.x;
        const bool valid = j < A.nwork;
        int cell = 0;
        if (valid) {
            // surface-chunk mode: the rank's hadrons are listed by chunk_own_kernel
            int s;
            int64_t ev, k;
            if (A.chunk) {
                const uint4 id = __ldg(&A.ident[j]);
                s = static_cast<int>(id.x);
                ev = id.y;
                k = id.z;
            } else {
                work_identity(A, sp_off, j, s, ev, k);
            }
            const DeviceSpecies p = sp[s];
            Task32 t;
            t.s = static_cast<uint16_t>(s);
            t.event = static_cast<uint32_t>(A.ev_begin + ev);
            t.draw = static_cast<uint32_t>(k);
            uint32_t w0, w1, w2, w3;
            philox_block(0u, t.draw, t.event, sample_stream_word3(s), key0, key1, w0, w1, w2, w3);
            cell = static_cast<int>(pick_cell(A, s, u53(w0, w1)));   // chunk mode: owned, hence >= 0
            t.cell = static_cast<uint32_t>(cell);
            // T and the chemical potentials come from the 16-byte-per-cell copy (15 MB at C4: it
            // stays in L2)
            const float4 tm = __ldg(A.thermo + cell);
            const float4 th = make_float4(tm.y, tm.z, tm.w, 0.f);
            MomSetup M;
            const bool ok = momentum_setup<true>(A.mt, p.mass, p.sign, tm.x, species_mu(p, 1, th, p.mass), M);
            t.m_term = M.m_term;
            t.cdf_max = M.cdf_max;
            t.tab_idx = ok ? static_cast<uint16_t>(M.tab | (M.idx_min << 3)) : TASK_RANGE_ERROR;
            if (!ok) my_range++;
            uint4 *dst = reinterpret_cast<uint4 *>(A.tasks_unsorted + j);
            const uint4 *src = reinterpret_cast<const uint4 *>(&t);
            dst[0] = src[0];
            dst[1] = src[1];
        }
        if (valid) atomicAdd(&hist[cell >> A.bucket_shift], 1u);
      }
      __syncthreads();
      // [block][segment]: the exclusive scan of this table in memory order is the write offset of
      // every (block, segment) run
      for (int b = threadIdx.x; b < A.nbucket; b += blockDim.x) {
          A.bucket_cnt[static_cast<int64_t>(b)*A.nseg + seg] = hist[b];
          hist[b] = 0u;
      }
      __syncthreads();
    }
    if (my_range) atomicAdd(&A.counters[3], my_range);
}

// K5b: after the exclusive scan of the [block][segment] histogram: every task to the region of its
// cell block.  A CTA takes a segment of PART_TILE consecutive tasks, orders them by block in shared
// memory (rank inside the block from a shared-memory counter, block bases from a scan of the
// counters) and copies the ordered tile out: consecutive threads write consecutive 16-byte halves,
// so every (block, segment) run of ~1 KB leaves as full, coalesced sectors at its scanned offset.
// No global atomics; inside a run the order depends on the timing of shared-memory atomics, the
// results do not.
__global__ void __launch_bounds__(PART_THREADS)
partition_kernel(const SamplerArgs A) {
    extern __shared__ __align__(16) unsigned char part_smem[];
    uint4 *stage = reinterpret_cast<uint4 *>(part_smem);                          // [PART_TILE][2]
    uint32_t *sslot = reinterpret_cast<uint32_t *>(stage + 2*PART_TILE);          // [PART_TILE] (chunk mode)
    __shared__ unsigned int cnt[MAX_BUCKETS];       // tasks of the segment per block
    __shared__ unsigned int lbase[MAX_BUCKETS + 1]; // exclusive scan of cnt
    __shared__ long long gbase[MAX_BUCKETS];        // global offset of the (block, segment) run
    for (int64_t seg = blockIdx.x; seg < A.nseg; seg += gridDim.x) {
        for (int b = threadIdx.x; b < A.nbucket; b += PART_THREADS) {
            cnt[b] = 0u;
            gbase[b] = static_cast<long long>(A.bucket_cnt[static_cast<int64_t>(b)*A.nseg + seg]);
        }
        __syncthreads();
        uint4 t0[PART_ITEMS], t1[PART_ITEMS];
        unsigned int rank[PART_ITEMS];
        const int64_t base = seg*PART_TILE;
        const int ntask = static_cast<int>(min(static_cast<int64_t>(PART_TILE), A.nwork - base));
#pragma unroll
        for (int i = 0; i < PART_ITEMS; i++) {
            const int l = i*PART_THREADS + threadIdx.x;
            if (l < ntask) {
                const uint4 *src = reinterpret_cast<const uint4 *>(A.tasks_unsorted + base + l);
                t0[i] = __ldg(src);
                t1[i] = __ldg(src + 1);
            }
        }
#pragma unroll
        for (int i = 0; i < PART_ITEMS; i++) {
            const int l = i*PART_THREADS + threadIdx.x;
            if (l < ntask) rank[i] = atomicAdd(&cnt[t1[i].x >> A.bucket_shift], 1u);
        }
        __syncthreads();
        if (threadIdx.x < 32) {
            // exclusive scan of at most MAX_BUCKETS (= 64) counters by one warp
            const int b0 = threadIdx.x, b1 = threadIdx.x + 32;
            const unsigned int c0 = (b0 < A.nbucket) ? cnt[b0] : 0u, c1 = (b1 < A.nbucket) ? cnt[b1] : 0u;
            unsigned int i0 = c0, i1 = c1;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const unsigned int x0 = __shfl_up_sync(0xffffffffu, i0, d);
                const unsigned int x1 = __shfl_up_sync(0xffffffffu, i1, d);
                if (threadIdx.x >= d) { i0 += x0; i1 += x1; }
            }
            const unsigned int tot0 = __shfl_sync(0xffffffffu, i0, 31);
            lbase[b0] = i0 - c0;
            lbase[b1] = tot0 + i1 - c1;
            if (threadIdx.x == 31) lbase[MAX_BUCKETS] = tot0 + i1;
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < PART_ITEMS; i++) {
            const int l = i*PART_THREADS + threadIdx.x;
            if (l < ntask) {
                const unsigned int p = lbase[t1[i].x >> A.bucket_shift] + rank[i];
                stage[2*p] = t0[i];
                stage[2*p + 1] = t1[i];
                if (A.chunk) sslot[p] = __ldg(&A.ident[base + l]).w;
            }
        }
        __syncthreads();
        // copy out: 16-byte half h of the ordered tile belongs to local task p = h/2, whose block
        // is read from the task itself (second half, word 0 = cell)
        for (int hidx = threadIdx.x; hidx < 2*ntask; hidx += PART_THREADS) {
            const int p = hidx >> 1;
            const unsigned int b = stage[2*p + 1].x >> A.bucket_shift;
            const int64_t pos = gbase[b] + (p - static_cast<int>(lbase[b]));
            reinterpret_cast<uint4 *>(A.tasks + pos)[hidx & 1] = stage[hidx];
            if (A.chunk && (hidx & 1) == 0) A.task_slot[pos] = sslot[p];
        }
        __syncthreads();
    }
}

// ---- bulk-async plumbing of the proposal kernel (mbarrier + cp.async.bulk, PTX ISA 8.x) ----------
__device__ __forceinline__ uint32_t smem_addr(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
                 :: "r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                 "selp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_addr(bar)), "r"(parity) : "memory");
    return ok != 0u;
}
// global -> shared bulk copy (bytes: multiple of 16, both addresses 16-byte aligned); completion
// is signalled on the mbarrier as transaction bytes
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_addr(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" :: "r"(smem_addr(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// what a lane of the proposal kernel carries
struct LaneState {
    uint32_t slot;          // index of the output record
    int s;
    int cell;
    BlockStream rng;
    MomSetup M;
    int tries;
    int total_tries;
    int qsign;      // +1 primary, -1 charge-conservation partner (flips B,S,Q)
    double eta_s;   // boost-invariant mode: eta_s of the primary, reused by its partner
};

// chunks 0..8 of a cell record -> the lane's shared-memory slot ([chunk][lane] layout: 16-byte
// accesses of a warp are conflict free)
__device__ __forceinline__ void lane_slot_fill(float4 *lc, int nthreads, int tid, const CellRec *rec) {
    const float4 *src = reinterpret_cast<const float4 *>(rec);
#pragma unroll
    for (int c = 0; c < CELLREC_SLOT_CHUNKS; c++) lc[c*nthreads + tid] = __ldg(src + c);
}

// Rare paths of the proposal kernel, kept out of line so that the hot loop stays small.
// (a) the reference's "impatience": after 4999 rejected tries a NEW cell is drawn (FSSW.cpp:1017-1018)
// (b) local charge conservation: the partner is sampled from the same cell with conjugate
//     quantum numbers (FSSW.cpp:1035-1048)
// The lane state crosses the call by value in a ColdIO block: a reference to LaneState itself
// would pin the whole state of the hot loop to local memory.
struct ColdIO {
    MomSetup M;
    uint32_t block, draw, event, slot;
    int s, cell, qsign, ok;
};

__device__ __noinline__ void lane_new_setup(const SamplerArgs *Ag, ColdIO *io, double mass, int sign,
                                            int B, int S, int Q, bool redraw_cell, uint32_t key0,
                                            uint32_t key1, float4 *lc, int nthreads, int tid) {
    // Ag: copy of the kernel arguments in global memory (taking the address of the by-value
    // kernel parameter would force a 1.2 KB per-thread stack copy)
    const SamplerArgs &A = *Ag;
    io->ok = 0;
    if (redraw_cell) {
        uint32_t w0, w1, w2, w3;
        philox_block(io->block++, io->draw, io->event, sample_stream_word3(io->s), key0, key1,
                     w0, w1, w2, w3);
        const int64_t c = pick_cell(A, io->s, u53(w0, w1));
        if (c < 0) {
            // surface-chunk mode: the new cell belongs to another rank (include/iss_cuda.h)
            atomicAdd(&A.counters[7], 1ull);
            float2 *dst = reinterpret_cast<float2 *>(A.out + io->slot);
#pragma unroll
            for (int q = 0; q < 5; q++) dst[q] = make_float2(0.f, 0.f);
            return;
        }
        io->cell = static_cast<int>(c);
        lane_slot_fill(lc, nthreads, tid, A.cellrec + io->cell);
    }
    const float4 th = __ldg(&A.cellrec[io->cell].th);      // T, muB, muS, muQ
    io->ok = momentum_setup(A.mt, mass, sign, th.x,
                            species_mu_bsq(io->qsign*B, io->qsign*S, io->qsign*Q, th.y, th.z, th.w, mass),
                            io->M) ? 1 : 0;
}

// SPEC selects a compile-time specialisation of the run-time mode flags (smaller and faster
// code for the common configurations); SPEC 0 is the generic kernel that handles every mode.
//   1: 3+1D, Chapman-Enskog (kind 21) shear + bulk + baryon diffusion, no charge pairing
//   2: 3+1D, Chapman-Enskog (kind 21) shear + bulk, no diffusion, no charge pairing
//   3: 3+1D, no delta f, no charge pairing
//   4: boost-invariant (2+1D), CE shear + bulk, no charge pairing (BASELINE.json configs[2])
template <int SPEC>
struct SpecMode {
    static constexpr bool generic = (SPEC == 0);
    static constexpr int shear = (SPEC == 1 || SPEC == 2 || SPEC == 4) ? 1 : 0;
    static constexpr int bulk = (SPEC == 1 || SPEC == 2 || SPEC == 4) ? 1 : 0;
    static constexpr int diff = (SPEC == 1) ? 1 : 0;
    static constexpr int kind = 21;
    static constexpr int neos = (SPEC == 3) ? -1 : 1;
    static constexpr int hydro_mode = (SPEC == 4) ? 1 : 2;
    static constexpr int lcc = 0;
};

// shared-memory copy of a regime-0 momentum table: [n][3] = CDF_0, CDF_1, CDF_2; the abscissa is
// recomputed as fma(i, dE, E_0), the expression build_momentum_table_kernel stores
__device__ __forceinline__ double table_F3(const double *tb, int i, double w1, double w0, double m_term) {
    const double c0 = tb[3*i], c1 = tb[3*i + 1], c2 = tb[3*i + 2];
    return c2 + w1*c1 + w0*c0 - m_term;
}

// MomentumSamplerBase::inverse_CDF (MomentumSamplerBase.cpp:61-93) on the shared-memory copy
__device__ __forceinline__ double inverse_cdf_smem(const double *tb, int n, double e0, double dE,
                                                   const MomSetup &M, double r) {
    const double w1 = 2.*M.mu_tilde;
    int lo = M.idx_min;
    int hi = n - 1;
    while (hi - lo > 1) {
        const int mid = static_cast<int>(static_cast<unsigned>(hi + lo) >> 1);  // (indices are >= 0)
        const double r_mid = table_F3(tb, mid, w1, M.w0, M.m_term);
        if (r < r_mid) hi = mid; else lo = mid;
    }
    // the values the bisection compared with at the final bracket (table_F3 is a pure function of the
    // index; hi = n - 1 was never probed: the reference starts from the cached maximum there)
    double r_min = table_F3(tb, lo, w1, M.w0, M.m_term);
    const double r_max = (hi == n - 1) ? M.cdf_max : table_F3(tb, hi, w1, M.w0, M.m_term);
    double E0 = __fma_rn(static_cast<double>(lo), dE, e0);
    if (E0 < M.a_min) {
        E0 = M.a_min;
        r_min = 0.;
    }
    const double Ehi = __fma_rn(static_cast<double>(hi), dE, e0);
    return E0 + (Ehi - E0)/fmax(1e-16, (r_max - r_min))*(r - r_min);
}

constexpr int RING_TASKS = 16;          // tasks per stage of a warp's task ring
constexpr int RING_STAGES = 2;
constexpr uint32_t RING_STAGE_BYTES = RING_TASKS*sizeof(Task32);

// K5c: persistent proposal kernel.  Every lane owns one hadron and repeats the reference's try
// (|p| proposal, direction, accept test) until it is accepted, then boosts, emits the record and
// immediately takes the next task: all 32 lanes of a warp stay busy whatever the acceptance rate.
// Tasks: warp w of the grid owns the stages w, w + W, w + 2W, ... (16 consecutive tasks each) of the
// cell-sorted list; lane 0 streams them two stages ahead into the warp's shared-memory ring with
// cp.async.bulk, completion on one mbarrier per stage, so that a task hand-over is a
// shared-memory read + one L1/L2-resident cell record (copied into the lane's slot with cp.async)
// instead of a chain of three dependent DRAM gathers.
template <int NTHREADS, int SPEC>
__global__ void __launch_bounds__(NTHREADS, 1)
propose_kernel(const SamplerArgs A, const SamplerArgs *Ag) {
    constexpr int SAMPLER_THREADS = NTHREADS;
    using SM = SpecMode<SPEC>;
    ModeFlags mode;
    mode.include_shear = SM::generic ? A.mode.include_shear : SM::shear;
    mode.include_bulk = SM::generic ? A.mode.include_bulk : SM::bulk;
    mode.include_diff = SM::generic ? A.mode.include_diff : SM::diff;
    mode.kind = SM::generic ? A.mode.kind : SM::kind;
    mode.neos = SM::generic ? A.mode.neos : SM::neos;
    const int hydro_mode = SM::generic ? A.hydro_mode : SM::hydro_mode;
    const int lcc = SM::generic ? A.lcc : SM::lcc;
    const uint32_t key0 = static_cast<uint32_t>(A.seed), key1 = static_cast<uint32_t>(A.seed >> 32);
    constexpr int NWARP = SAMPLER_THREADS/32;

    // shared memory: per-lane cell slots, the two regime-0 momentum tables without their abscissa
    // (boson 24 KB, fermion 48 KB: they serve every species with (m - mu)/T < 30 and take the
    // 11-step bisection of each proposal off the L1/LSU gather path), task rings, species table
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4 *lc = reinterpret_cast<float4 *>(smem_raw);                  // [9][SAMPLER_THREADS]
    Task32 *ring_all = reinterpret_cast<Task32 *>(lc + CELLREC_SLOT_CHUNKS*SAMPLER_THREADS);
    uint64_t *bar_all = reinterpret_cast<uint64_t *>(ring_all + NWARP*RING_STAGES*RING_TASKS);
    PropSpecies *sp = reinterpret_cast<PropSpecies *>(bar_all + NWARP*RING_STAGES);
    double *sm_boson = reinterpret_cast<double *>(sp + A.ns);
    double *sm_fermion = sm_boson + 3*A.mt[0].n;
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    Task32 *ring = ring_all + warp*RING_STAGES*RING_TASKS;
    uint64_t *bar = bar_all + warp*RING_STAGES;
    // slot chunks: 0 da | 1 pa | 2 pb | 3 u4 | 4 pos | 5 tz, inv_T | 6 inv_dsig, shear | 7 cb, c1 |
    // 8 inv_kappa, prefq
#define LC4(c) lc[(c)*SAMPLER_THREADS + tid]
#define LCD(c, k) (reinterpret_cast<const double *>(&lc[(c)*SAMPLER_THREADS + tid])[k])

    for (int i = tid; i < A.ns; i += SAMPLER_THREADS) {
        const DeviceSpecies d = A.species[i];
        PropSpecies q;
        q.mass = d.mass;
        q.mass2 = d.mass2;
        q.pid = d.pid;
        q.baryon = d.baryon;
        q.strange = d.strange;
        q.charge = d.charge;
        q.sign = d.sign;
        q.pad = 0;
        sp[i] = q;
    }
    for (int i = tid; i < 3*A.mt[0].n; i += SAMPLER_THREADS)
        sm_boson[i] = __ldg(A.mt[0].data + 4*(i/3) + 1 + i%3);
    for (int i = tid; i < 3*A.mt[3].n; i += SAMPLER_THREADS)
        sm_fermion[i] = __ldg(A.mt[3].data + 4*(i/3) + 1 + i%3);
    if (lane == 0) {
#pragma unroll
        for (int b = 0; b < RING_STAGES; b++) mbar_init(&bar[b], 1u);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const unsigned full = 0xffffffffu;
    const unsigned lt_mask = (1u << lane) - 1u;
    // stages of this warp: g = gw + k W, k = 0, 1, ...; stage k lives in ring buffer k & 1
    // (a batch holds fewer than 2^32 tasks: stage indices fit 32 bits)
    const int W = static_cast<int>(gridDim.x)*NWARP;
    const int gw = static_cast<int>(blockIdx.x)*NWARP + warp;
    const int nstage = static_cast<int>((A.nwork + RING_TASKS - 1)/RING_TASKS);
    auto issue_stage = [&](int buf, int g) {
        // (lane 0) the ring buffer was last read through the generic proxy
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_arrive_expect_tx(&bar[buf], RING_STAGE_BYTES);
        bulk_g2s(ring + buf*RING_TASKS, A.tasks + static_cast<int64_t>(g)*RING_TASKS, RING_STAGE_BYTES,
                 &bar[buf]);
    };
    if (lane == 0) {
#pragma unroll
        for (int b = 0; b < RING_STAGES; b++)
            if (gw + b*W < nstage) issue_stage(b, gw + b*W);
    }
    // tasks in stage g (0 past the end of the list)
    const uint32_t nwork32 = static_cast<uint32_t>(A.nwork);
    auto stage_tasks = [&](int g) -> int {
        const uint32_t left = nwork32 - static_cast<uint32_t>(g)*RING_TASKS;
        return left >= RING_TASKS ? RING_TASKS : static_cast<int>(left);
    };
    int k_stage = 0;            // stages this warp has used up
    int cursor = 0;             // tasks of the current stage handed out
    bool stage_ready = false;   // the current stage's bytes have landed
    bool busy = false;
    LaneState L;
    L.qsign = 1;
    unsigned long long my_tries = 0, my_redraws = 0, my_range = 0, my_giveup = 0;

    for (;;) {
        // ------------------------------------------------------------ hand tasks to idle lanes
        // (a) fetch: idle lanes copy their task out of the ring (the loop turns again only when a
        // stage runs out in the middle of a hand-over); (b) set-up, once per round for all of them
        unsigned need_mask = __ballot_sync(full, !busy);
        bool fresh = false;
        double2 t0 = make_double2(0., 0.);
        uint4 t1 = make_uint4(0u, 0u, 0u, 0u);
        uint32_t task_pos = 0;
        while (need_mask != 0u) {
            const int g = gw + k_stage*W;
            if (g >= nstage) break;
            const int buf = k_stage & (RING_STAGES - 1);
            if (!stage_ready) {
                const uint32_t parity = static_cast<uint32_t>(k_stage/RING_STAGES) & 1u;
                while (!mbar_try_wait(&bar[buf], parity)) { }
                stage_ready = true;
            }
            const int n_cur = stage_tasks(g);
            const int take = min(__popc(need_mask), n_cur - cursor);
            const int rank = __popc(need_mask & lt_mask);
            if (!busy && !fresh && rank < take) {
                const int i = buf*RING_TASKS + cursor + rank;
                t0 = *reinterpret_cast<const double2 *>(&ring[i]);          // m_term, cdf_max
                t1 = *(reinterpret_cast<const uint4 *>(&ring[i]) + 1);      // cell, event, draw, s | tab
                task_pos = static_cast<uint32_t>(g)*RING_TASKS + cursor + rank;
                fresh = true;
            }
            cursor += take;
            if (cursor == n_cur) {
                // every lane's reads of this ring buffer are done: refill it two stages ahead
                __syncwarp();
                if (lane == 0 && g + RING_STAGES*W < nstage) issue_stage(buf, g + RING_STAGES*W);
                k_stage++;
                cursor = 0;
                stage_ready = false;
            }
            need_mask = __ballot_sync(full, !busy && !fresh);
        }
        if (fresh) {
            L.cell = static_cast<int>(t1.x);
            L.rng.event = t1.y;
            L.rng.draw = t1.z;
            L.rng.block = 1u;               // block 0 chose the cell (setup_kernel)
            L.s = static_cast<int>(t1.w & 0xFFFFu);
            const uint32_t tab_idx = t1.w >> 16;
            L.qsign = 1;
            L.tries = 1;
            L.total_tries = 0;
            if (tab_idx != TASK_RANGE_ERROR) {
                const float4 *src = reinterpret_cast<const float4 *>(A.cellrec + L.cell);
#pragma unroll
                for (int c = 0; c < CELLREC_SLOT_CHUNKS; c++) cp_async16(&LC4(c), src + c);
                cp_async_commit();
                const float4 th = __ldg(src + CELLREC_SLOT_CHUNKS);     // T, muB, muS, muQ
                const PropSpecies p = sp[L.s];
                if (A.chunk) {
                    L.slot = __ldg(&A.task_slot[task_pos]);
                } else {
                    // output slot: hadron `draw` of (event, species); pairs under charge conservation
                    const int mult = (lcc == 1 && p.charge > 0) ? 2 : 1;
                    const int64_t ev = static_cast<int64_t>(t1.y) - A.ev_begin;
                    L.slot = static_cast<uint32_t>(__ldg(&A.off_out[ev*A.ns + L.s])
                                                   + static_cast<int64_t>(t1.z)*mult);
                }
                momentum_restore(p.mass, th.x,
                                 species_mu_bsq(p.baryon, p.strange, p.charge, th.y, th.z, th.w, p.mass),
                                 t0.x, t0.y, static_cast<int>(tab_idx & 7u),
                                 static_cast<int>(tab_idx >> 3), L.M);
                busy = true;
            }
            // TASK_RANGE_ERROR: momentum table range error, counted by setup_kernel; no record
        }
        if (!__any_sync(full, busy)) break;     // idle lanes are left only when the stages ran out

        // ------------------------------------------------------------ one try per busy lane
        if (busy) {
            cp_async_wait_all();                // the lane's slot (no-op after the first try)
            const PropSpecies p = sp[L.s];
            const double mass = p.mass;
            const int sign = p.sign;
            const MomentumTable &mt = A.mt[L.M.tab];
            // the rare paths exchange the lane state by value (ColdIO)
            auto cold_setup = [&](bool redraw) -> bool {
                ColdIO io;
                io.block = L.rng.block; io.draw = L.rng.draw; io.event = L.rng.event;
                io.slot = L.slot; io.s = L.s; io.cell = L.cell; io.qsign = L.qsign;
                lane_new_setup(Ag, &io, mass, sign, p.baryon, p.strange, p.charge, redraw, key0, key1,
                               lc, SAMPLER_THREADS, tid);
                L.rng.block = io.block;
                L.cell = io.cell;
                L.M = io.M;
                L.tries = 1;
                return io.ok != 0;
            };
            // |p| proposal (MomentumSamplerBase.cpp:48-56); an inner rejection restarts the try
            uint32_t pw0, pw1, pw2, pw3;
            philox_block(L.rng.block++, L.rng.draw, L.rng.event, sample_stream_word3(L.s), key0, key1,
                         pw0, pw1, pw2, pw3);
            const double r = u53(pw0, pw1)*L.M.cdf_max;
            // regime-0 tables (every species with (m - mu)/T < 30): ONE code path for bosons and
            // fermions, the lanes of a warp differ in the table base only (cell-sorted tasks mix
            // the species inside a warp)
            double Et;
            if ((L.M.tab == 0 || L.M.tab == 3) && A.mt_smem)
                Et = inverse_cdf_smem(L.M.tab == 0 ? sm_boson : sm_fermion, mt.n, mt.e0, mt.de_build, L.M, r);
            else Et = inverse_cdf<false>(mt.data, mt.n, L.M, r);
            const double E_sample = L.M.T*Et + L.M.mu;
            const double p_mag = sqrt(E_sample*E_sample - p.mass2);
            // (p/E)/(1 - m^2/(2E^2)) = 2 p E/(2 E^2 - m^2): one division instead of three
            const double accept_ratio = 2.*p_mag*E_sample/(2.*E_sample*E_sample - p.mass2);
            const double u_inner = u32(pw2);
            if (!(u_inner > accept_ratio)) {
                my_tries++;
                L.total_tries++;
                // FSSW.cpp:1880-1945
                uint32_t qw0, qw1, qw2, qw3;
                philox_block(L.rng.block++, L.rng.draw, L.rng.event, sample_stream_word3(L.s), key0,
                             key1, qw0, qw1, qw2, qw3);
                // phi = 2 pi u: sin/cos through sincospi(2u) (no large-argument reduction path)
                const double u_phi = u32(pw3);
                const double cos_theta = 2.*u32(qw0) - 1.;
                const double sin_theta = sqrt(1. - cos_theta*cos_theta);
                const double pT = p_mag*sin_theta;
                double sphi, cphi;
                sincospi(2.*u_phi, &sphi, &cphi);
                const double px = pT*cphi;
                const double py = pT*sphi;
                const double p0 = sqrt(p.mass2 + p_mag*p_mag);
                const double pz = p_mag*cos_theta;
                const float4 da = LC4(0);
                const double pdsigma = p0*da.x + px*da.y + py*da.z + pz*da.w;
                const double inv_p0 = 1.0/p0;
                // p.dsigma/(p0 (|dsigma0| + |dsigma_vec|)), FSSW.cpp:1939
                double fact1 = pdsigma*inv_p0*LCD(6, 0);
                fact1 = fmax(0., fmin(1., fact1));
                // the accept uniform is drawn whatever delta f is; since fact2 <= 1, u >= fact1
                // already decides "reject" and the delta-f evaluation is skipped
                const double u_acc = u32(qw1);
                double accept_prob = 0.;
                if (u_acc < fact1) {
                    const double inv_T = LCD(5, 1);
                    const double f0 = 1./(exp((p0 - L.M.mu)*inv_T) + sign);
                    const double stat = 1. - sign*f0;
                    double delta_f = 0.;
                    if (mode.include_shear | mode.include_bulk | mode.include_diff) {
                        const int B = L.qsign*p.baryon, S = L.qsign*p.strange, Q = L.qsign*p.charge;
                        if (mode.include_shear == 1) {
                            const float4 pa = LC4(1);           // pixx, pixy, pixz, piyy
                            const float4 pb = LC4(2);           // piyz, qx, qy, qz
                            const double Wfactor = (px*px*pa.x + 2.*px*py*pa.y + 2.*px*pz*pa.z
                                                    + py*py*pa.w + 2.*py*pz*pb.x
                                                    + pz*pz*(-pa.x - pa.w));
                            const double sh = LCD(6, 1);
                            if (mode.neos == 1) delta_f += stat*Wfactor*sh*inv_p0*inv_T;
                            else delta_f += stat*Wfactor*sh;
                        }
                        if (mode.include_bulk == 1) {
                            // FSSW::get_deltaf_bulk (FSSW.cpp:1795-1849); kinds 0,2,3,4: bulkPi = 0
                            if (mode.kind == 21 || mode.kind == 1) {
                                // -(1 -/+ f0) c0 (m^2/(3 T p0) - c1 p0/T) Pi
                                delta_f += (-stat*LCD(7, 0)*inv_T
                                            *(p.mass2*(1./3.)*inv_p0 - LCD(7, 1)*p0));
                            } else if (mode.kind == 11 || mode.kind == 20) {
                                const double *__restrict__ co =
                                    A.cellcoef + static_cast<int64_t>(L.cell)*COEF_STRIDE;
                                const double bulkPi = __ldg(reinterpret_cast<const float4 *>(
                                    A.cells + static_cast<int64_t>(L.cell)*CELL_STRIDE) + 4).w;
                                if (mode.kind == 11) {
                                    delta_f += stat*bulkPi*(__ldg(&co[0])*p.mass2 + __ldg(&co[1])*B*p0
                                                            + __ldg(&co[2])*p0*p0);
                                } else {
                                    delta_f += stat*bulkPi*(p.mass2*__ldg(&co[2])
                                                            + p0*(B*__ldg(&co[3]) + S*__ldg(&co[4])
                                                                  + Q*__ldg(&co[5]))
                                                            + p0*p0*(__ldg(&co[1]) - __ldg(&co[2])));
                                }
                            }
                        }
                        if (mode.include_diff == 1) {
                            const float4 pb = LC4(2);           // piyz, qx, qy, qz
                            const double qmufactor = -px*pb.y - py*pb.z - pz*pb.w;
                            delta_f += stat*(LCD(8, 1) - B*inv_p0)*qmufactor*LCD(8, 0);
                        }
                    }
                    double fact2 = (1. + delta_f)/2.;
                    fact2 = fmax(0., fmin(1., fact2));
                    accept_prob = fact1*fact2;
                }
                if (u_acc < accept_prob) {
                    // ---- accepted: boost to the lab frame and emit (FSSW.cpp:1946-1960, 1969-1996)
                    const float4 pos = LC4(4);          // tau, x, y, eta
                    const float4 u4 = LC4(3);           // ut, ux, uy, uz
                    const float pl0 = static_cast<float>(p0), pl1 = static_cast<float>(px),
                                pl2 = static_cast<float>(py), pl3 = static_cast<float>(pz);
                    double p_dot_u = 0.;
                    p_dot_u += __fmul_rn(pl1, u4.y);
                    p_dot_u += __fmul_rn(pl2, u4.z);
                    p_dot_u += __fmul_rn(pl3, u4.w);
                    const float up1 = __fadd_rn(u4.x, 1.f);
                    const double fac = p_dot_u/up1 + pl0;
                    const float lab1 = static_cast<float>(pl1 + fac*u4.y);
                    const float lab2 = static_cast<float>(pl2 + fac*u4.z);
                    const float lab3 = static_cast<float>(pl3 + fac*u4.w);
                    iss_hadron hd;
                    hd.pid = (L.qsign > 0) ? p.pid : -p.pid;
                    hd.mass = static_cast<float>(mass);
                    hd.x = pos.y;
                    hd.y = pos.z;
                    const double pT_lab = sqrt(static_cast<double>(__fadd_rn(__fmul_rn(lab1, lab1),
                                                                             __fmul_rn(lab2, lab2))));
                    const double mT = sqrt(pT_lab*pT_lab + mass*mass);
                    hd.px = lab1;
                    hd.py = lab2;
                    if (hydro_mode == 2) {
                        // eta_s = cell eta: y = asinh(pz/mT) - eta + eta, p_z = mT sinh(y) = pLab[3],
                        // px = pT cos(atan2(py,px)) = pLab[1] up to FP64 rounding; t,z precomputed.
                        const float4 tzi = LC4(5);          // t, z of the cell (| inv_T)
                        const double pzl = lab3;
                        hd.pz = lab3;
                        hd.E = static_cast<float>(sqrt(mT*mT + pzl*pzl));
                        hd.t = tzi.x;
                        hd.z = tzi.y;
                    } else {
                        // boost-invariant: y ~ U(y_LB, y_RB), eta_s = y - (y - eta_s) (FSSW.cpp:1024-1029);
                        // the charge-conservation partner keeps the primary's eta_s (FSSW.cpp:1045-1047)
                        const double y_minus_eta = asinh(lab3/mT) - pos.w;
                        if (L.qsign > 0) {
                            uint32_t w0, w1, w2, w3;
                            philox_block(L.rng.block++, L.rng.draw, L.rng.event,
                                         sample_stream_word3(L.s), key0, key1, w0, w1, w2, w3);
                            const double rap = A.y_LB + (A.y_RB - A.y_LB)*u32(w0);
                            L.eta_s = rap - y_minus_eta;
                        }
                        const double eta_s = L.eta_s;
                        const double rapidity_y = y_minus_eta + eta_s;
                        hd.pz = static_cast<float>(mT*sinh(rapidity_y));
                        hd.E = static_cast<float>(mT*cosh(rapidity_y));
                        hd.z = static_cast<float>(pos.x*sinh(eta_s));
                        hd.t = static_cast<float>(pos.x*cosh(eta_s));
                    }
                    // 40-byte record as five 8-byte stores (slots are 8-byte aligned)
                    float2 *dst = reinterpret_cast<float2 *>(A.out + L.slot);
                    dst[0] = make_float2(__int_as_float(hd.pid), hd.mass);
                    dst[1] = make_float2(hd.E, hd.px);
                    dst[2] = make_float2(hd.py, hd.pz);
                    dst[3] = make_float2(hd.t, hd.x);
                    dst[4] = make_float2(hd.y, hd.z);
                    if (A.trace_cell) {
                        A.trace_cell[L.slot] = L.cell;
                        A.trace_tries[L.slot] = L.total_tries;
                    }
                    busy = false;
                    if (lcc == 1 && L.qsign > 0 && p.charge > 0) {
                        // partner with conjugate quantum numbers from the SAME cell
                        L.qsign = -1;
                        L.slot += 1;
                        L.total_tries = 0;
                        if (cold_setup(false)) busy = true;
                        else my_range++;
                    }
                } else {
                    L.tries++;
                    if (L.tries >= MAX_IMPATIENCE) {
                        if (L.total_tries > MAX_TRIES_PER_HADRON) {
                            // The reference would loop forever on a (cell, species) pair that can
                            // never be accepted (FSSW.cpp:977-1018); a kernel must not.  The slot
                            // gets a null record (pid 0), the offender is recorded for the error
                            // message and the call reports ISS_ERR_RANGE.
                            float2 *dst = reinterpret_cast<float2 *>(A.out + L.slot);
#pragma unroll
                            for (int q = 0; q < 5; q++) dst[q] = make_float2(0.f, 0.f);
                            if (my_giveup == 0) {
                                A.giveup_info[0] = static_cast<unsigned long long>(L.cell + A.cell_begin);
                                A.giveup_info[1] = static_cast<unsigned long long>(L.s);
                            }
                            my_giveup++;
                            busy = false;
                        } else if (L.qsign > 0) {
                            // the reference's "impatience" (FSSW.cpp:1017-1018 with status == 0)
                            my_redraws++;
                            if (!cold_setup(true)) {
                                my_range++;
                                busy = false;
                            }
                        } else {
                            // partner sampling never re-picks the cell (do-while at FSSW.cpp:1039-1044)
                            L.tries = 1;
                        }
                    }
                }
            }
        }
    }
#undef LC4
#undef LCD
    // warp-aggregated counters
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        my_tries += __shfl_down_sync(full, my_tries, d);
        my_redraws += __shfl_down_sync(full, my_redraws, d);
        my_range += __shfl_down_sync(full, my_range, d);
        my_giveup += __shfl_down_sync(full, my_giveup, d);
    }
    if (lane == 0) {
        if (my_giveup) atomicAdd(&A.counters[6], my_giveup);
        if (my_tries) atomicAdd(&A.counters[1], my_tries);
        if (my_redraws) atomicAdd(&A.counters[2], my_redraws);
        if (my_range) atomicAdd(&A.counters[3], my_range);
    }
}

// Unit-level entry for the |p| sampler alone (rows M of SURVEY.md section 8(a)): n draws of
// MomentumSamplerShell::Sample_a_momentum(m, T, mu, sign) (MomentumSamplerShell.cpp:24-48), draw i
// from stream (seed; SAMPLE, species 0, event i >> 20, draw i & 0xFFFFF), one proposal block per
// iteration of the inner accept loop -- the protocol of the proposal kernel.
__global__ void momentum_unit_kernel(const MomentumTable *__restrict__ mts, double m, double T,
                                     double mu, int sign, int64_t n, uint64_t seed,
                                     double *__restrict__ out, int *__restrict__ range_error) {
    const int64_t i = static_cast<int64_t>(blockIdx.x)*blockDim.x + threadIdx.x;
    if (i >= n) return;
    MomSetup M;
    if (!momentum_setup(mts, m, sign, T, mu, M)) {
        *range_error = 1;
        return;
    }
    const MomentumTable &mt = mts[M.tab];
    const uint32_t key0 = static_cast<uint32_t>(seed), key1 = static_cast<uint32_t>(seed >> 32);
    uint32_t block = 0;
    double p_mag, ratio;
    uint32_t w0, w1, w2, w3;
    do {
        philox_block(block++, static_cast<uint32_t>(i & 0xFFFFF), static_cast<uint32_t>(i >> 20),
                     sample_stream_word3(0), key0, key1, w0, w1, w2, w3);
        const double r = u53(w0, w1)*M.cdf_max;
        const double Et = inverse_cdf<false>(mt.data, mt.n, M, r);
        const double E = M.T*Et + M.mu;
        p_mag = sqrt(E*E - m*m);
        ratio = (p_mag/E)/(1. - m*m/(2.*E*E));
    } while (u32(w2) > ratio);
    out[i] = p_mag;
}

int run_momentum_unit(iss_handle *h, double m, double T, double mu, int sign, int64_t n,
                      uint64_t seed, double *out_host) {
    int rc = ensure_momentum_tables(h);
    if (rc) return rc;
    MomentumTable *d_mt = nullptr;
    double *d_out = nullptr;
    int *d_err = nullptr;
    ISS_CUDA_TRY(h, cudaMalloc(&d_mt, sizeof(MomentumTable)*6));
    ISS_CUDA_TRY(h, cudaMalloc(&d_out, sizeof(double)*n));
    ISS_CUDA_TRY(h, cudaMalloc(&d_err, sizeof(int)));
    ISS_CUDA_TRY(h, cudaMemsetAsync(d_err, 0, sizeof(int), h->stream));
    ISS_CUDA_TRY(h, cudaMemcpyAsync(d_mt, h->momtab, sizeof(MomentumTable)*6, cudaMemcpyHostToDevice,
                                    h->stream));
    momentum_unit_kernel<<<static_cast<unsigned>((n + 127)/128), 128, 0, h->stream>>>(
        d_mt, m, T, mu, sign, n, seed, d_out, d_err); ISS_LAUNCHED(h);
    int err = 0;
    cudaError_t e1 = cudaMemcpyAsync(out_host, d_out, sizeof(double)*n, cudaMemcpyDeviceToHost, h->stream);
    cudaError_t e2 = cudaMemcpyAsync(&err, d_err, sizeof(int), cudaMemcpyDeviceToHost, h->stream);
    cudaError_t e3 = cudaStreamSynchronize(h->stream);
    cudaFree(d_mt); cudaFree(d_out); cudaFree(d_err);
    if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess)
        ISS_FAIL(h, ISS_ERR_CUDA, "momentum unit kernel failed");
    if (err) ISS_FAIL(h, ISS_ERR_RANGE, "[MomentumSampler] out of range (m/T - mu/T outside the tables)");
    return ISS_OK;
}

// ---------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------
int run_multiplicities(iss_handle *h, uint64_t seed, int64_t nev) {
    const int ns = h->nspecies;
    const int64_t n = nev*ns;
    int rc;
    if (n + 1 > h->mult_cap || !h->d_mult) {
        if (h->d_mult) cudaFree(h->d_mult);
        if (h->d_off_out) cudaFree(h->d_off_out);
        if (h->d_off_work) cudaFree(h->d_off_work);
        h->d_mult = h->d_off_out = h->d_off_work = nullptr;
        h->mult_cap = n + 1 + n/8;
        ISS_CUDA_TRY(h, cudaMalloc(&h->d_mult, sizeof(int64_t)*h->mult_cap));
        ISS_CUDA_TRY(h, cudaMalloc(&h->d_off_out, sizeof(int64_t)*h->mult_cap));
        ISS_CUDA_TRY(h, cudaMalloc(&h->d_off_work, sizeof(int64_t)*h->mult_cap));
    }
    rc = ensure_capacity(h, &h->d_event_off, &h->event_off_cap, nev + 1);
    if (rc) return rc;

    // Poisson parameters per species (host, once per yields computation)
    const iss_options &o = h->opt;
    h->h_lambda.resize(ns);
    h->h_pmode.resize(ns);
    for (int s = 0; s < ns; s++) {
        double dN = h->h_total[s];
        if (o.hydro_mode != 2) dN = (o.y_RB - o.y_LB)*dN;     // FSSW.cpp:953-958
        h->h_lambda[s] = dN;
        double pm = 1.0;
        if (dN >= 1e-15) {
            const double m = floor(dN);
            pm = exp(m*log(dN) - dN - lgamma(m + 1.0));
        }
        h->h_pmode[s] = pm;
    }
    if (!h->d_lambda) {
        ISS_CUDA_TRY(h, cudaMalloc(&h->d_lambda, sizeof(double)*MAX_SPECIES));
        ISS_CUDA_TRY(h, cudaMalloc(&h->d_pmode, sizeof(double)*MAX_SPECIES));
    }
    if (!h->lambda_on_device) {
        ISS_CUDA_TRY(h, cudaMemcpyAsync(h->d_lambda, h->h_lambda.data(), sizeof(double)*ns,
                                        cudaMemcpyHostToDevice, h->stream));
        ISS_CUDA_TRY(h, cudaMemcpyAsync(h->d_pmode, h->h_pmode.data(), sizeof(double)*ns,
                                        cudaMemcpyHostToDevice, h->stream));
        // the host vectors are pageable: the copies are complete when the calls return
        h->lambda_on_device = true;
    }

    ScopedTimer t(h, ISS_T_MULT);
    // out_count -> d_off_out (scanned in place), work_count -> d_off_work (in place)
    ISS_CUDA_TRY(h, cudaMemsetAsync(h->d_off_out + n, 0, sizeof(int64_t), h->stream));
    ISS_CUDA_TRY(h, cudaMemsetAsync(h->d_off_work + n, 0, sizeof(int64_t), h->stream));
    multiplicity_kernel<<<static_cast<unsigned>((n + 255)/256), 256, 0, h->stream>>>(
        h->d_lambda, h->d_pmode, h->d_species, ns, nev, h->ev_begin, seed,
        o.dN_dy_sampling_model, o.dN_dy_sampling_para1, o.local_charge_conservation, h->d_mult,
        h->d_off_out,
        h->d_off_work); ISS_LAUNCHED(h);
    ISS_CUDA_TRY(h, cudaGetLastError());
    return ISS_OK;
}

// Surface-chunk mode: how many hadrons of the batch are this rank's.  d_off_work holds the scanned
// work offsets of the WHOLE batch (identical on every rank); afterwards d_own holds the exclusive
// prefix over the (species, event) pairs of the owned draws, d_ownmask the ballot words, d_off_out
// the UNSCANNED per (event, species) output counts of this rank; their number is posted to mail
// slot 9 (read after the next stream synchronisation).
static int chunk_count_work(iss_handle *h, const SamplerArgs &A, int64_t nev, int64_t total_work, int nsm) {
    const int64_t npair = nev*h->nspecies;
    int rc = ensure_capacity(h, &h->d_own, &h->own_cap, npair + 1);
    if (rc) return rc;
    rc = ensure_capacity(h, &h->d_ownmask, &h->ownmask_cap, (total_work >> 5) + npair + 2);
    if (rc) return rc;
    ISS_CUDA_TRY(h, cudaMemsetAsync(h->d_own + npair, 0, sizeof(int64_t), h->stream));
    const int64_t grid = std::min<int64_t>((npair + 7)/8, static_cast<int64_t>(nsm)*8);
    chunk_own_kernel<<<static_cast<unsigned>(grid), 256, 0, h->stream>>>(A, h->d_own, h->d_off_out,
                                                                        h->d_ownmask); ISS_LAUNCHED(h);
    ISS_CUDA_TRY(h, cudaGetLastError());
    rc = device_exclusive_scan_i64(h, h->d_own, h->d_own, npair, nullptr);
    if (rc) return rc;
    return mail_post(h, h->d_own + npair, 1, 9);
}

}  // namespace iss
#define ISS_LEGACY_WITH_SAMPLER
#include "legacy.cuh"
namespace iss {

// per-cell records of the proposal kernel (after the yield kernel has left the delta-f coefficients)
int build_cellrec(iss_handle *h) {
    ISS_ENSURE(h, h->d_cellrec, h->cellrec_bytes, sizeof(CellRec)*h->ncell);
    const iss_options &o = h->opt;
    ModeFlags mode;
    mode.include_shear = o.include_deltaf_shear;
    mode.include_bulk = o.include_deltaf_bulk;
    mode.include_diff = o.include_deltaf_diffusion;
    mode.kind = o.bulk_deltaf_kind;
    mode.neos = (o.bulk_deltaf_kind == 21) ? 1 : (o.bulk_deltaf_kind == 20 ? 0 : -1);
    build_cellrec_kernel<<<static_cast<unsigned>((h->ncell + 127)/128), 128, 0, h->stream>>>(
        h->d_cells, h->d_cellcoef, h->ncell, mode, static_cast<CellRec *>(h->d_cellrec)); ISS_LAUNCHED(h);
    ISS_CUDA_TRY(h, cudaGetLastError());
    h->cellrec_valid = true;
    return ISS_OK;
}

int run_sampler(iss_handle *h, uint64_t seed, int64_t nev, int64_t /*unused*/) {
    const int ns = h->nspecies;
    const int64_t n = nev*ns;
    int rc = h->legacy ? ISS_OK : ensure_momentum_tables(h);
    if (rc) return rc;
    int dev = 0, nsm = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);

    const iss_options &o = h->opt;
    SamplerArgs A;
    A.cells = h->d_cells;
    A.thermo = h->d_thermo;
    A.cellcoef = h->d_cellcoef;
    A.ncell = h->ncell;
    A.ncell_pad = h->ncell_pad;
    A.cdf = h->d_cdf;
    A.cdflev = h->d_cdflev;
    A.nlev = h->nlev;
    for (int k = 0; k < 8; k++) A.lev_off[k] = h->lev_off[k];
    A.lev_stride = h->lev_stride;
    A.total = h->d_total;
    A.guide = (h->guide_M > 0 && !h->chunk) ? static_cast<const int2 *>(h->d_guide) : nullptr;
    A.guide_M = h->guide_M;
    A.species = h->d_species;
    A.ns = ns;
    A.nev = nev;
    A.ev_begin = h->ev_begin;
    A.off_work = h->d_off_work;
    A.off_out = h->d_off_out;
    for (int r = 0; r < 6; r++) A.mt[r] = h->momtab[r];
    A.mode.include_shear = o.include_deltaf_shear;
    A.mode.include_bulk = o.include_deltaf_bulk;
    A.mode.include_diff = o.include_deltaf_diffusion;
    A.mode.kind = o.bulk_deltaf_kind;
    A.mode.neos = (o.bulk_deltaf_kind == 21) ? 1 : (o.bulk_deltaf_kind == 20 ? 0 : -1);
    A.hydro_mode = o.hydro_mode;
    A.lcc = o.local_charge_conservation;
    A.y_LB = o.y_LB;
    A.y_RB = o.y_RB;
    A.seed = seed;
    A.trace_cell = nullptr;
    A.trace_tries = nullptr;
    A.chunk = h->chunk ? 1 : 0;
    A.g_nlev = h->g_nlev;
    A.cdflev_g = h->d_cdflev_g;
    for (int k = 0; k < 8; k++) A.g_lev_off[k] = h->g_lev_off[k];
    A.g_lev_stride = h->g_lev_stride;
    A.cell_begin = h->chunk_cell_begin;
    A.g_ncell = h->g_ncell;
    A.blk_begin = h->chunk_cell_begin/ISS_CHUNK_ALIGN;
    // the last chunk also owns the partly filled block at the end of the surface
    A.blk_end = (h->chunk_cell_begin + h->ncell >= h->g_ncell)
                    ? (int64_t(1) << 62) : (h->chunk_cell_begin + h->ncell)/ISS_CHUNK_ALIGN;
    A.ident = nullptr;
    A.hints = nullptr;
    A.tasks = nullptr;
    A.task_slot = nullptr;
    A.tasks_unsorted = nullptr;
    A.bucket_cnt = nullptr;
    A.bucket_shift = 0;
    A.nbucket = 1;
    A.nseg = 0;
    A.cellrec = nullptr;
    A.giveup_info = nullptr;
    A.mt_smem = 0;
    A.out = nullptr;
    A.counters = nullptr;
    A.nwork = 0;

    int64_t total_out = 0, total_work = 0, nhint = 0;
    {
        ScopedTimer t(h, ISS_T_MULT);
        rc = device_exclusive_scan_i64(h, h->d_off_work, h->d_off_work, n, &total_work);
        if (rc) return rc;
        nhint = (total_work + SETUP_THREADS - 1)/SETUP_THREADS;
        {
            const size_t need = sizeof(int2)*static_cast<size_t>(nhint > 0 ? nhint : 1);
            if (need > h->hints_bytes || !h->d_hints) {
                if (h->d_hints) cudaFree(h->d_hints);
                h->d_hints = nullptr;
                h->hints_bytes = need + need/8 + 4096;
                ISS_CUDA_TRY(h, cudaMalloc(&h->d_hints, h->hints_bytes));
            }
            A.hints = static_cast<const int2 *>(h->d_hints);
        }
        const bool chunk_work = h->chunk && total_work > 0;
        if (chunk_work) {
            rc = chunk_count_work(h, A, nev, total_work, nsm);
            if (rc) return rc;
        }
        rc = device_exclusive_scan_i64(h, h->d_off_out, h->d_off_out, n, &total_out);
        if (rc) return rc;
        // (the scan's synchronisation also delivered the number of owned hadrons)
        A.nwork = chunk_work ? static_cast<int64_t>(*reinterpret_cast<volatile unsigned long long *>(h->h_mail + 9))
                             : total_work;
        rc = ensure_mapped_event_offsets(h, nev + 1);
        if (rc) return rc;
        event_offset_kernel<<<static_cast<unsigned>((nev + 1 + 255)/256), 256, 0, h->stream>>>(
            h->d_off_out, ns, nev, h->d_event_off, h->d_evoff_mapped); ISS_LAUNCHED(h);
        ISS_CUDA_TRY(h, cudaGetLastError());
    }
    {
        // output buffer of this batch: the one whose previous device->host copy (if any) is
        // ordered before this kernel
        const int b = h->cur_buf;
        if (h->copy_pending[b]) {
            ISS_CUDA_TRY(h, cudaStreamWaitEvent(h->stream, h->copy_done[b], 0));
            h->copy_pending[b] = false;
        }
        rc = ensure_capacity(h, &h->d_hadbuf[b], &h->hadbuf_cap[b], total_out);
        if (rc) return rc;
        h->d_hadrons = h->d_hadbuf[b];
        h->hadron_cap = h->hadbuf_cap[b];
    }
    if (!h->d_counters) ISS_CUDA_TRY(h, cudaMalloc(&h->d_counters, sizeof(unsigned long long)*N_COUNTERS));
    ISS_CUDA_TRY(h, cudaMemsetAsync(h->d_counters, 0, sizeof(unsigned long long)*N_COUNTERS, h->stream));
    h->n_hadrons = total_out;
    h->n_primaries = total_out;
    if (A.nwork == 0) return ISS_OK;

    A.out = h->d_hadrons;
    A.counters = h->d_counters;
    if (h->trace) {
        if (total_out > h->trace_cap || !h->d_trace) {
            if (h->d_trace) cudaFree(h->d_trace);
            h->d_trace = nullptr;
            h->trace_cap = total_out + 1024;
            ISS_CUDA_TRY(h, cudaMalloc(&h->d_trace, sizeof(int32_t)*2*h->trace_cap));
        }
        A.trace_cell = h->d_trace;
        A.trace_tries = h->d_trace + h->trace_cap;
    }

    if (h->legacy) {
        // legacy sampler (MC_sampling = 2): one persistent kernel, lanes take work items one by one
        LegacyArgs G;
        rc = legacy_args(h, G);
        if (rc) return rc;
        {
            ScopedTimer t(h, ISS_T_SETUP);
            work_hint_kernel<<<static_cast<unsigned>((nhint + 127)/128), 128, 0, h->stream>>>(
                h->d_off_work, ns, nev, total_work, static_cast<int2 *>(h->d_hints), nhint); ISS_LAUNCHED(h);
        }
        const size_t smem_l = sizeof(DeviceSpecies)*ns + sizeof(int64_t)*(ns + 1)
                              + sizeof(float)*LEGACY_LANE_STRIDE*LEGACY_THREADS;
        int64_t grid_l = static_cast<int64_t>(nsm)*4;
        const int64_t useful = (A.nwork + LEGACY_THREADS - 1)/LEGACY_THREADS;
        if (grid_l > useful) grid_l = useful;
        if (smem_l > 48*1024)
            ISS_CUDA_TRY(h, cudaFuncSetAttribute(legacy_sample_kernel,
                                                 cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                 static_cast<int>(smem_l)));
        {
            ScopedTimer t(h, ISS_T_SAMPLE);
            legacy_sample_kernel<<<static_cast<unsigned>(grid_l), LEGACY_THREADS, smem_l, h->stream>>>(A, G);
            ISS_LAUNCHED(h);
        }
        ISS_CUDA_TRY(h, cudaGetLastError());
        return ISS_OK;
    }

    // cell-sorted task list of the batch (padded to whole ring stages: the proposal kernel copies
    // whole stages) and the scratch of the counting sort
    if (total_out >= (int64_t(1) << 32) || A.nwork >= (int64_t(1) << 32))
        ISS_FAIL(h, ISS_ERR_ARG, "a batch must hold fewer than 2^32 hadrons (sample fewer events per call)");
    const int64_t nstage = (A.nwork + RING_TASKS - 1)/RING_TASKS;
    {
        const size_t need = sizeof(Task32)*static_cast<size_t>(nstage*RING_TASKS);
        if (need > h->tasks_bytes || !h->d_tasks) {
            if (h->d_tasks) cudaFree(h->d_tasks);
            if (h->d_tasks_unsorted) cudaFree(h->d_tasks_unsorted);
            h->d_tasks = nullptr;
            h->d_tasks_unsorted = nullptr;
            h->tasks_bytes = need + need/8 + 4096;
            ISS_CUDA_TRY(h, cudaMalloc(&h->d_tasks, h->tasks_bytes));
            ISS_CUDA_TRY(h, cudaMalloc(&h->d_tasks_unsorted, h->tasks_bytes));
        }
        // cell blocks of the partition: at most MAX_BUCKETS (the 2048 tasks of a segment leave in
        // runs of >= 32 per block = 1 KB), at least 256 cells each; the cell records of the one or two
        // blocks in flight (2.6 MB each at C4) stay in L2
        A.bucket_shift = 8;
        while (((h->ncell - 1) >> A.bucket_shift) + 1 > MAX_BUCKETS) A.bucket_shift++;
        A.nbucket = static_cast<int>(((h->ncell - 1) >> A.bucket_shift) + 1);
        A.nseg = (A.nwork + PART_TILE - 1)/PART_TILE;
        // (capacities with head-room: the batch size fluctuates from call to call, and a
        // cudaFree + cudaMalloc pair synchronises the device)
        if (sizeof(unsigned long long)*(static_cast<size_t>(A.nbucket)*A.nseg + 2) > h->cellcnt_bytes)
            ISS_ENSURE(h, h->d_cellcnt, h->cellcnt_bytes,
                       sizeof(unsigned long long)*(static_cast<size_t>(A.nbucket)*(A.nseg + A.nseg/8 + 16) + 2));
        A.tasks = static_cast<Task32 *>(h->d_tasks);
        A.tasks_unsorted = static_cast<Task32 *>(h->d_tasks_unsorted);
        A.bucket_cnt = h->d_cellcnt;
        if (h->chunk) {     // output slots travel through the sort (they do not follow from the draw index)
            const size_t slot_need = sizeof(uint32_t)*static_cast<size_t>(nstage*RING_TASKS);
            if (slot_need > h->task_slot_bytes)
                ISS_ENSURE(h, h->d_task_slot, h->task_slot_bytes, slot_need + slot_need/8 + 4096);
            A.task_slot = h->d_task_slot;
            // identities of the rank's hadrons: two int64 of d_wlist per hadron
            rc = ensure_capacity(h, &h->d_wlist, &h->wlist_cap, 2*A.nwork + 2);
            if (rc) return rc;
            A.ident = reinterpret_cast<const uint4 *>(h->d_wlist);
        }
    }
    if (!h->cellrec_valid) {
        rc = build_cellrec(h);
        if (rc) return rc;
    }
    A.cellrec = static_cast<const CellRec *>(h->d_cellrec);
    A.giveup_info = h->d_counters + 8;
    A.mt_smem = (h->momtab[0].generated && h->momtab[3].generated) ? 1 : 0;
    if (!h->d_sampler_args) ISS_CUDA_TRY(h, cudaMalloc(&h->d_sampler_args, sizeof(SamplerArgs)));
    ISS_CUDA_TRY(h, cudaMemcpyAsync(h->d_sampler_args, &A, sizeof(SamplerArgs), cudaMemcpyHostToDevice,
                                    h->stream));
    {
        ScopedTimer t(h, ISS_T_SETUP);
        if (!h->chunk) {
            work_hint_kernel<<<static_cast<unsigned>((nhint + 127)/128), 128, 0, h->stream>>>(
                h->d_off_work, ns, nev, total_work, static_cast<int2 *>(h->d_hints), nhint); ISS_LAUNCHED(h);
        } else {
            // (d_off_out is scanned by now: the output slots are known)
            const int64_t npair = nev*ns;
            const int64_t cgrid = std::min<int64_t>((npair + 7)/8, static_cast<int64_t>(nsm)*8);
            chunk_fill_kernel<<<static_cast<unsigned>(cgrid), 256, 0, h->stream>>>(
                A, h->d_own, h->d_off_out, h->d_ownmask, reinterpret_cast<uint4 *>(h->d_wlist)); ISS_LAUNCHED(h);
        }
        // partition of the batch's hadrons by cell block: histogram per segment, exclusive scan, scatter
        const size_t smem_setup = sizeof(DeviceSpecies)*ns + sizeof(int64_t)*(ns + 1)
                                  + sizeof(unsigned int)*A.nbucket;
        const int64_t sgrid = std::min<int64_t>(A.nseg, static_cast<int64_t>(nsm)*32);
        setup_kernel<<<static_cast<unsigned>(sgrid), SETUP_THREADS, smem_setup, h->stream>>>(A); ISS_LAUNCHED(h);
        const int64_t ncnt = static_cast<int64_t>(A.nbucket)*A.nseg;
        ISS_CUDA_TRY(h, cudaMemsetAsync(h->d_cellcnt + ncnt, 0, sizeof(unsigned long long), h->stream));
        rc = device_exclusive_scan_i64(h, reinterpret_cast<const int64_t *>(h->d_cellcnt),
                                       reinterpret_cast<int64_t *>(h->d_cellcnt), ncnt, nullptr);
        if (rc) return rc;
        const int64_t pgrid = std::min<int64_t>(A.nseg, static_cast<int64_t>(nsm)*16);
        const size_t smem_part = (sizeof(uint4)*2 + sizeof(uint32_t))*PART_TILE;
        ISS_CUDA_TRY(h, cudaFuncSetAttribute(partition_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             static_cast<int>(smem_part)));
        partition_kernel<<<static_cast<unsigned>(pgrid), PART_THREADS, smem_part, h->stream>>>(A);
        ISS_LAUNCHED(h);
    }
    ISS_CUDA_TRY(h, cudaGetLastError());

    int spec = 0;
    if (A.hydro_mode == 2 && A.lcc != 1) {
        const bool ce = (A.mode.kind == 21 && A.mode.include_shear == 1 && A.mode.include_bulk == 1);
        if (ce && A.mode.include_diff == 1) spec = 1;
        else if (ce && A.mode.include_diff != 1) spec = 2;
        else if (A.mode.include_shear != 1 && A.mode.include_bulk != 1 && A.mode.include_diff != 1
                 && A.mode.kind != 20 && A.mode.kind != 21) spec = 3;
    }
    if (A.hydro_mode != 2 && A.lcc != 1 && A.mode.kind == 21 && A.mode.include_shear == 1
        && A.mode.include_bulk == 1 && A.mode.include_diff != 1) spec = 4;
    static int force_generic = -1;
    if (force_generic < 0) {
        const char *e = getenv("ISS_SAMPLER_GENERIC");
        force_generic = (e && atoi(e) == 1) ? 1 : 0;
    }
    if (force_generic) spec = 0;
    // CTA size: 768 threads (80 registers each); ISS_SAMPLER_THREADS=640 / 512 select the tuning
    // variants of the CE + diffusion specialisation (96 / 128 registers)
    int nthreads = SAMPLER_THREADS;
    void (*kern)(const SamplerArgs, const SamplerArgs *) =
        spec == 1 ? propose_kernel<SAMPLER_THREADS, 1> : spec == 2 ? propose_kernel<SAMPLER_THREADS, 2>
        : spec == 3 ? propose_kernel<SAMPLER_THREADS, 3> : spec == 4 ? propose_kernel<SAMPLER_THREADS, 4>
        : propose_kernel<SAMPLER_THREADS, 0>;
    static int tune_threads = -1;
    if (tune_threads < 0) {
        const char *e = getenv("ISS_SAMPLER_THREADS");
        tune_threads = e ? atoi(e) : 0;
    }
    if (spec == 1 && tune_threads == 640) { kern = propose_kernel<640, 1>; nthreads = 640; }
    if (spec == 1 && tune_threads == 512) { kern = propose_kernel<512, 1>; nthreads = 512; }
    const int nwarp = nthreads/32;
    const size_t smem = sizeof(float4)*CELLREC_SLOT_CHUNKS*nthreads
                        + sizeof(Task32)*nwarp*RING_STAGES*RING_TASKS
                        + sizeof(uint64_t)*nwarp*RING_STAGES + sizeof(PropSpecies)*ns
                        + sizeof(double)*3*(A.mt[0].n + A.mt[3].n);
    {
        int smem_max = 0;
        cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
        if (static_cast<int>(smem) > smem_max)
            ISS_FAIL(h, ISS_ERR_ARG, "too many species for the proposal kernel's shared-memory species table");
    }
    ISS_CUDA_TRY(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem)));
    int64_t grid = nsm;         // persistent: one CTA per SM
    const int64_t max_useful = (A.nwork + nthreads - 1)/nthreads;
    if (grid > max_useful) grid = max_useful;
    {
        ScopedTimer t(h, ISS_T_SAMPLE);
        kern<<<static_cast<unsigned>(grid), nthreads, smem, h->stream>>>(
            A, static_cast<const SamplerArgs *>(h->d_sampler_args)); ISS_LAUNCHED(h);
    }
    ISS_CUDA_TRY(h, cudaGetLastError());
    return ISS_OK;
}

}  // namespace iss
