// spectra.cu -- smooth Cooper-Frye spectra dN/(pT dpT dphi dy).
//
// Replaces EmissionFunctionArray::calculate_dN_pTdpTdphidy (emissionfunction.cpp:624-829) with
// its helpers getbulkvisCoefficients (:3625-3762), get_deltaf_bulk (:4156-4186) and
// get_deltaf_qmu_coeff (:3788-3823).  The reference loops pT x phi x cells x (y - eta_s) per
// species with ~60 FP64 operations and one exp per point: a dense reduction, FP64-pipe bound.
//
// Mapping: one thread owns one (pT, phi) point of one species and sums over the cells of one
// cell chunk in registers, so no reduction tree is needed; the chunk partials are added in chunk
// order by a second kernel (deterministic).  All threads of a CTA walk the same cells, so
// everything that depends on the cell (and on the cell and the y - eta_s point) only is computed
// once per CTA into shared memory and read back as broadcasts:
//   p.u     = mT (ch u0 - sh u3)            - (px u1 + py u2)            = mT A[c][k] - Bv
//   p.dsig  = mT (ch da0 + sh da3/tau)      + (px da1 + py da2)          = mT C[c][k] + Dv
//   W       = mT^2 (ch^2 pi00 - 2 ch sh pi03 + sh^2 pi33)
//             + mT (ch (-2)(px pi01 + py pi02) + sh 2 (px pi13 + py pi23))
//             + px^2 pi11 + 2 px py pi12 + py^2 pi22                     = mT^2 E[c][k] + mT (ch F + sh G) + H
//   p.q     = mT (ch q0 - sh q3)            - (px q1 + py q2)            = mT Q[c][k] - R
// with ch = cosh(-(y - eta_s)), sh = sinh(-(y - eta_s)), p^tau = mT ch, p^eta = mT sh.
// This regroups the reference's sums (differences ~1e-15 relative; the parity tests allow 1e-10).
#include "iss_internal.cuh"
#include "coefficients.cuh"
#include "legacy_bulk.cuh"

#include <cmath>
#include <type_traits>

namespace iss {

namespace {

constexpr int SPEC_THREADS = 256;
#ifndef SPEC_UNROLL
#define SPEC_UNROLL 3
#endif
constexpr int K_UNROLL = SPEC_UNROLL;    // y - eta_s points interleaved per thread
constexpr int CELL_TILE = 8;            // cells staged per shared-memory tile
constexpr int MAX_NY = 128;             // y - eta_s points
constexpr int REC = 36;                 // doubles per cell record (species independent)
enum {
    R_U0 = 0, R_U1, R_U2, R_U3, R_DA0, R_DA1, R_DA2, R_DA3T, R_INVT, R_TAU, R_SHEAR,
    R_PI00, R_PI01, R_PI02, R_PI03, R_PI11, R_PI12, R_PI13, R_PI22, R_PI23, R_PI33,
    R_BULKPI, R_C0, R_C1, R_C2, R_Q0, R_Q1, R_Q2, R_Q3, R_INVKAPPA, R_PREFQ,
    R_MUB, R_MUS, R_MUQ, R_T, R_SPARE
};
// per-(CTA, cell) scalars in shared memory
enum {
    S_U1 = 0, S_U2, S_DA1, S_DA2, S_PI01, S_PI02, S_PI13, S_PI23, S_PI11, S_PI12, S_PI22,
    S_Q1, S_Q2, S_MU, S_INVT, S_T, S_SHEAR, S_BULKPI, S_C0, S_C1, S_C2, S_INVKAPPA, S_PREFQ, S_TAUFAC,
    S_COUNT
};


// exp() and reciprocal of the occupation number, written out so that every constant is a
// constant-bank operand of the DFMA that uses it (the library exp() re-materialises its 64-bit
// coefficients through uniform-register moves on every call: ~24 extra issue slots per point)
// and without special-case branches.  [0] 16/ln2, [1] 1.5 * 2^52, [2] -(ln2/16) high part,
// [3] -(ln2/16) low part, [4..10] 1/6! ... 1/0!
__constant__ double c_exp[11] = {
    23.083120654223414519, 6755399441055744.0, -4.33216987730702385306e-02,
    -1.19263433079411731251e-11,
    1.0/720.0, 1.0/120.0, 1.0/24.0, 1.0/6.0, 0.5, 1.0, 1.0};

// 1/d for a normal, finite d: hardware seed (20+ bits) and two Newton steps, no slow path
__device__ __forceinline__ double fast_rcp(double d) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
    double e = fma(-d, y, 1.0);
    y = fma(y, e, y);
    e = fma(-d, y, 1.0);
    return fma(y, e, y);
}

// f0 = 1/(exp(x) + sign) and exp(x) f0 = 1 - sign f0.  exp: x = (16 n + j) ln2/16 + r,
// |r| <= ln2/32, exp(x) = 2^n 2^(j/16) p(r) with a degree-6 Taylor polynomial (remainder
// < 5e-16) and a 16-entry table of 2^(j/16) in shared memory.  2^n goes through the exponent
// field with n clamped to [-1022, 1010]: beyond +-700 the reference's f0 is below 1e-304 of the
// table's scale (its exp overflows to inf and f0 becomes 0 there).
__device__ __forceinline__ double occupation(double x, double sign, const double *s_pow2) {
    const double t = fma(x, c_exp[0], c_exp[1]);
    const int n16 = __double2loint(t);
    const double tn = t - c_exp[1];
    double r = fma(tn, c_exp[2], x);
    r = fma(tn, c_exp[3], r);
    double p = c_exp[4];
#pragma unroll
    for (int i = 5; i < 11; i++) p = fma(p, r, c_exp[i]);
    const int n = min(max(n16 >> 4, -1022), 1010);
    const double scale = __hiloint2double((n + 1023) << 20, 0);
    return fast_rcp(fma(p*s_pow2[n16 & 15], scale, sign));
}

struct SpectraArgs {
    const float *lab;           // [ncell][ISS_LAB_NFIELD]
    double *rec;                // [ncell][REC]
    int64_t ncell;
    iss_spectra_options opt;
    CoefTables tabs;            // kappa only
    // momentum grid and y - eta_s table (device): pT[npT], cos phi[nphi], sin phi[nphi],
    // cosh[ny], sinh[ny], weight[ny]
    const double *pT, *cphi, *sphi, *ch, *sh, *wy;
    int npT, nphi, ny;
    const DeviceSpecies *species;
    int ns;
    int64_t chunk;              // cells per chunk
    int nchunk;
    double *part;               // [nchunk][ns][npT*nphi][2] (sum, max)
    double *out;                // [2][ns][npT*nphi]
};

// Species-independent part of the per-cell set-up (emissionfunction.cpp:688-757), once per call.
__global__ void __launch_bounds__(128)
spectra_cell_kernel(const SpectraArgs A) {
    const int64_t cell = static_cast<int64_t>(blockIdx.x)*blockDim.x + threadIdx.x;
    if (cell >= A.ncell) return;
    const float4 *cr = reinterpret_cast<const float4 *>(A.lab + cell*ISS_LAB_NFIELD);
    float f[ISS_LAB_NFIELD];
#pragma unroll
    for (int q = 0; q < ISS_LAB_NFIELD/4; q++) {
        const float4 v = __ldg(cr + q);
        f[4*q] = v.x; f[4*q + 1] = v.y; f[4*q + 2] = v.z; f[4*q + 3] = v.w;
    }
    double *r = A.rec + cell*REC;
    const double T = f[ISS_L_T], P = f[ISS_L_P], E = f[ISS_L_E], tau = f[ISS_L_TAU];
    r[R_U0] = f[ISS_L_U0]; r[R_U1] = f[ISS_L_U1]; r[R_U2] = f[ISS_L_U2]; r[R_U3] = f[ISS_L_U3];
    r[R_DA0] = f[ISS_L_DA0]; r[R_DA1] = f[ISS_L_DA1]; r[R_DA2] = f[ISS_L_DA2];
    r[R_DA3T] = static_cast<double>(f[ISS_L_DA3])/tau;
    r[R_INVT] = 1.0/T;
    r[R_T] = T;
    r[R_TAU] = tau;
    r[R_SHEAR] = A.opt.include_deltaf_shear ? 1.0/(2.0*T*T*(E + P)) : 0.0;
    r[R_PI00] = f[ISS_L_PI00]; r[R_PI01] = f[ISS_L_PI01]; r[R_PI02] = f[ISS_L_PI02];
    r[R_PI03] = f[ISS_L_PI03]; r[R_PI11] = f[ISS_L_PI11]; r[R_PI12] = f[ISS_L_PI12];
    r[R_PI13] = f[ISS_L_PI13]; r[R_PI22] = f[ISS_L_PI22]; r[R_PI23] = f[ISS_L_PI23];
    r[R_PI33] = f[ISS_L_PI33];
    double bulkPi = 0., c0 = 0., c1 = 0.;
    if (A.opt.include_deltaf_bulk == 1) {
        const int kind = A.opt.bulk_deltaf_kind;
        if (kind == 0) {
            bulkPi = f[ISS_L_BULKPI];       // coefficients stay zero (emissionfunction.cpp:728-730)
        } else {
            bulkPi = static_cast<double>(f[ISS_L_BULKPI])/HBARC;
            if (kind >= 1 && kind <= 4) {
                legacy_bulk_poly(kind, T, c0, c1);
            }
        }
    }
    r[R_BULKPI] = bulkPi; r[R_C0] = c0; r[R_C1] = c1; r[R_C2] = 0.;
    double inv_kappa = 0., pref_q = 0.;
    if (A.opt.include_deltaf_diffusion == 1) {
        inv_kappa = 1.0/coef_kappa(A.tabs, T, static_cast<double>(f[ISS_L_MUB]));
        pref_q = static_cast<double>(f[ISS_L_BN])/(E + P);
    }
    r[R_Q0] = f[ISS_L_Q0]; r[R_Q1] = f[ISS_L_Q1]; r[R_Q2] = f[ISS_L_Q2]; r[R_Q3] = f[ISS_L_Q3];
    r[R_INVKAPPA] = inv_kappa; r[R_PREFQ] = pref_q;
    r[R_MUB] = f[ISS_L_MUB]; r[R_MUS] = f[ISS_L_MUS]; r[R_MUQ] = f[ISS_L_MUQ];
    r[R_SPARE] = 0.;
}

// BULK: 0 none/kind 0 (zero coefficients), 1..4 the reference's kinds; DIFF: diffusion delta f;
// WANT_MAX: also the maximum over (cell, y - eta_s) the reference keeps for its MC_sampling = 3
template <int BULK, bool DIFF, bool WANT_MAX>
__global__ void __launch_bounds__(SPEC_THREADS, 2)
spectra_kernel(const SpectraArgs A) {
    __shared__ double s_ch[MAX_NY], s_sh[MAX_NY], s_wy[MAX_NY];
    __shared__ double s_pow2[16];
    __shared__ double2 s_ac[CELL_TILE][MAX_NY];     // A, C
    __shared__ double2 s_eq[CELL_TILE][MAX_NY];     // E, Q
    __shared__ double s_sc[CELL_TILE][S_COUNT];

    const int tile = blockIdx.x, s = blockIdx.y, chunk = blockIdx.z;
    const int npt = A.npT*A.nphi;
    const int m = tile*SPEC_THREADS + threadIdx.x;
    const bool live = m < npt;
    const DeviceSpecies sp = A.species[s];
    const double mass = sp.mass;
    const double sign = sp.sign;
    const int ny = A.ny;
    if (threadIdx.x < 16) s_pow2[threadIdx.x] = exp2(threadIdx.x*(1.0/16.0));
    for (int k = threadIdx.x; k < ny; k += blockDim.x) {
        s_ch[k] = A.ch[k];
        s_sh[k] = A.sh[k];
        s_wy[k] = A.wy[k];
    }
    double pT = 0., px = 0., py = 0.;
    if (live) {
        const int i = m/A.nphi, j = m - i*A.nphi;
        pT = A.pT[i];
        px = pT*A.cphi[j];
        py = pT*A.sphi[j];
    }
    const double mT = sqrt(mass*mass + pT*pT);
    const double mT2 = mT*mT;
    const double pxx = px*px, pxy2 = 2.0*px*py, pyy = py*py;
    const double mass2 = mass*mass;
    const double baryon = sp.baryon;
    const double ratio_max = A.opt.deltaf_max_ratio;
    const bool restrict_df = A.opt.restrict_deltaf == 1;
    const bool pos_only = A.opt.use_pos_dN_only != 0;
    // 1/(8 pi^3)/hbarc^3 * degeneracy (emissionfunction.cpp:644, 796)
    const double pref = 1.0/(8.0*(M_PI*M_PI*M_PI))/HBARC/HBARC/HBARC*static_cast<double>(sp.gspin);

    double sum = 0., vmax = 0.;
    const int64_t c_begin = static_cast<int64_t>(chunk)*A.chunk;
    const int64_t c_end = min(c_begin + A.chunk, A.ncell);
    for (int64_t c0 = c_begin; c0 < c_end; c0 += CELL_TILE) {
        const int nc = static_cast<int>(min(static_cast<int64_t>(CELL_TILE), c_end - c0));
        __syncthreads();        // previous tile consumed (and the k tables written)
        // ---- per-cell scalars
        for (int q = threadIdx.x; q < nc; q += blockDim.x) {
            const double *r = A.rec + (c0 + q)*REC;
            double *o = s_sc[q];
            o[S_U1] = r[R_U1]; o[S_U2] = r[R_U2]; o[S_DA1] = r[R_DA1]; o[S_DA2] = r[R_DA2];
            o[S_PI01] = -2.0*r[R_PI01]; o[S_PI02] = -2.0*r[R_PI02];
            o[S_PI13] = 2.0*r[R_PI13]; o[S_PI23] = 2.0*r[R_PI23];
            o[S_PI11] = r[R_PI11]; o[S_PI12] = r[R_PI12]; o[S_PI22] = r[R_PI22];
            o[S_Q1] = r[R_Q1]; o[S_Q2] = r[R_Q2];
            // int * float products summed in float, as the reference's
            // `double mu = baryon*surf->muB + strange*surf->muS + charge*surf->muQ` (:700)
            const float muB = static_cast<float>(r[R_MUB]), muS = static_cast<float>(r[R_MUS]),
                        muQ = static_cast<float>(r[R_MUQ]);
            const float mu = __fadd_rn(__fadd_rn(__fmul_rn(static_cast<float>(sp.baryon), muB),
                                                 __fmul_rn(static_cast<float>(sp.strange), muS)),
                                       __fmul_rn(static_cast<float>(sp.charge), muQ));
            o[S_MU] = static_cast<double>(mu)*r[R_INVT];       // mu/T
            o[S_INVT] = r[R_INVT];
            o[S_T] = r[R_T];
            o[S_SHEAR] = r[R_SHEAR];
            o[S_BULKPI] = r[R_BULKPI]; o[S_C0] = r[R_C0]; o[S_C1] = r[R_C1]; o[S_C2] = r[R_C2];
            o[S_INVKAPPA] = r[R_INVKAPPA]; o[S_PREFQ] = r[R_PREFQ];
            o[S_TAUFAC] = pref*r[R_TAU];
        }
        // ---- per-(cell, y - eta_s) terms
        for (int e = threadIdx.x; e < nc*ny; e += blockDim.x) {
            const int q = e/ny, k = e - q*ny;
            const double *r = A.rec + (c0 + q)*REC;
            const double ch = s_ch[k], sh = s_sh[k];
            double2 ac, eq;
            ac.x = ch*r[R_U0] - sh*r[R_U3];
            ac.y = pref*r[R_TAU]*(ch*r[R_DA0] + sh*r[R_DA3T]);    // with prefactor * g * tau
            eq.x = r[R_SHEAR]*(ch*ch*r[R_PI00] - 2.0*ch*sh*r[R_PI03] + sh*sh*r[R_PI33]);   // x shear prefactor
            eq.y = ch*r[R_Q0] - sh*r[R_Q3];
            s_ac[q][k] = ac;
            s_eq[q][k] = eq;
        }
        __syncthreads();
        if (!live) continue;
        for (int q = 0; q < nc; q++) {
            const double *o = s_sc[q];
            const double Bv = px*o[S_U1] + py*o[S_U2];
            const double taufac = o[S_TAUFAC];
            const double Dv = taufac*(px*o[S_DA1] + py*o[S_DA2]);
            const double shear = o[S_SHEAR];
            // W x shear prefactor = mT^2 E'[k] + ch F + sh G + H
            const double F = shear*mT*(px*o[S_PI01] + py*o[S_PI02]);
            const double G = shear*mT*(px*o[S_PI13] + py*o[S_PI23]);
            const double H = shear*(pxx*o[S_PI11] + pxy2*o[S_PI12] + pyy*o[S_PI22]);
            const double Rq = px*o[S_Q1] + py*o[S_Q2];
            const double mu_T = o[S_MU], invT = o[S_INVT];
            const double bulkPi = o[S_BULKPI], bc0 = o[S_C0], bc1 = o[S_C1];
            const double inv_kappa = o[S_INVKAPPA], pref_q = o[S_PREFQ];
            const double Tc = o[S_T];
            const double m2_3T = mass2*invT*(1.0/3.0);      // (m/T)^2/(3 E/T) = m^2/(3 T) / p.u
            // MODE 0: no restriction of delta f; 2: resize = min(1, ratio/(|df| + 1e-10)) at
            // every point, branch-free (a data-dependent branch inside the loop keeps the
            // scheduler from interleaving the unrolled points and costs more than the reciprocal;
            // on viscous surfaces the restriction bites too often for a detect-and-redo scheme)
            auto cell_loop = [&](auto mode_tag, double &csum, double &cmax) -> bool {
                constexpr int MODE = decltype(mode_tag)::value;
                bool exceed = false;
                csum = 0.;
                cmax = 0.;
#pragma unroll K_UNROLL
                for (int k = 0; k < ny; k++) {
                    const double2 ac = s_ac[q][k];
                    const double2 eq = s_eq[q][k];
                    const double pdotu = fma(mT, ac.x, -Bv);
                    const double f0 = occupation(fma(pdotu, invT, -mu_T), sign, s_pow2);
                    const double pdsigma = fma(mT, ac.y, Dv);           // x prefactor * g * tau
                    const double one_m = fma(-sign, f0, 1.0);
                    double df = one_m*fma(mT2, eq.x, fma(s_ch[k], F, fma(s_sh[k], G, H)));
                    if (BULK != 0 || DIFF) {
                        const double EoT = pdotu*invT;
                        double inv_pdotu = 0.;
                        if (BULK == 1 || BULK == 4 || DIFF) inv_pdotu = fast_rcp(pdotu);
                        if (BULK == 1) {
                            df += -one_m*bc0*(m2_3T*inv_pdotu - bc1*EoT)*bulkPi;
                        } else if (BULK == 2) {
                            df += -one_m*bulkPi*(-bc0 + bc1*EoT);
                        } else if (BULK == 3) {
                            df += -one_m*bulkPi*rsqrt(EoT)*(-bc0 + bc1*EoT);
                        } else if (BULK == 4) {
                            df += -one_m*bulkPi*(bc0 - bc1*Tc*inv_pdotu);
                        }
                        if (DIFF) df += one_m*(pref_q - baryon*inv_pdotu)*(mT*eq.y - Rq)*inv_kappa;
                    }
                    if (MODE == 2) {
                        const double size = fabs(df) + 1e-10;
                        const double resized = df*(ratio_max*fast_rcp(size));
                        df = size > ratio_max ? resized : df;
                    }
                    const double fp = f0*pdsigma;
                    const double result = fma(fp, df, fp);
                    if (!(pos_only && result < 0.)) {
                        csum = fma(result, s_wy[k], csum);
                        if (WANT_MAX) cmax = fmax(cmax, result);
                    }
                }
                return exceed;
            };
            double csum, cmax;
            if (restrict_df) {
                cell_loop(std::integral_constant<int, 2>(), csum, cmax);
            } else {
                cell_loop(std::integral_constant<int, 0>(), csum, cmax);
            }
            sum += csum;
            if (WANT_MAX) vmax = fmax(vmax, cmax);
        }
    }
    if (live) {
        double2 *o = reinterpret_cast<double2 *>(A.part)
                     + (static_cast<int64_t>(chunk)*A.ns + s)*npt + m;
        *o = make_double2(sum, vmax);
    }
}

// adds the chunk partials in chunk order
__global__ void __launch_bounds__(256)
spectra_reduce_kernel(const SpectraArgs A) {
    const int64_t n = static_cast<int64_t>(A.ns)*A.npT*A.nphi;
    const int64_t i = static_cast<int64_t>(blockIdx.x)*blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double2 *p = reinterpret_cast<const double2 *>(A.part);
    double sum = 0., vmax = 0.;
    for (int c = 0; c < A.nchunk; c++) {
        const double2 v = p[static_cast<int64_t>(c)*n + i];
        sum += v.x;
        vmax = fmax(vmax, v.y);
    }
    A.out[i] = sum;
    A.out[n + i] = vmax;
}

template <int BULK>
void launch_spectra(const SpectraArgs &A, dim3 grid, cudaStream_t st, bool want_max) {
    const bool diff = A.opt.include_deltaf_diffusion == 1;
    if (diff && want_max) spectra_kernel<BULK, true, true><<<grid, SPEC_THREADS, 0, st>>>(A);
    else if (diff) spectra_kernel<BULK, true, false><<<grid, SPEC_THREADS, 0, st>>>(A);
    else if (want_max) spectra_kernel<BULK, false, true><<<grid, SPEC_THREADS, 0, st>>>(A);
    else spectra_kernel<BULK, false, false><<<grid, SPEC_THREADS, 0, st>>>(A);
}

}  // namespace

}  // namespace iss

using namespace iss;

extern "C" {

int iss_cuda_upload_surface_lab(iss_handle *h, const float *cells, int64_t ncell) {
    if (!h || !cells || ncell <= 0) return ISS_ERR_ARG;
    cudaSetDevice(h->device);
    ISS_ENSURE(h, h->d_lab, h->lab_bytes, sizeof(float)*ISS_LAB_NFIELD*ncell);
    ISS_CUDA_TRY(h, cudaMemcpyAsync(h->d_lab, cells, sizeof(float)*ISS_LAB_NFIELD*ncell,
                                    cudaMemcpyHostToDevice, h->stream));
    ISS_CUDA_TRY(h, cudaStreamSynchronize(h->stream));      // the caller's buffer may be reused
    h->nlab = ncell;
    if (h->legacy) {        // yields of the legacy sampler belonged to the previous cells
        h->have_yields = false;
        h->have_batch = false;
    }
    return ISS_OK;
}

int iss_cuda_spectra(iss_handle *h, const iss_spectra_options *opt, const iss_species *species,
                     int32_t nspecies, const double *pT, int32_t npT, const double *phi,
                     int32_t nphi, const double *y_minus_eta, const double *y_weight, int32_t ny,
                     double *dN, double *dN_max) {
    if (!h || !opt || !species || !pT || !phi || !y_minus_eta || !y_weight || !dN) return ISS_ERR_ARG;
    if (nspecies <= 0 || nspecies > 65535 || npT <= 0 || nphi <= 0 || ny <= 0 || ny > MAX_NY)
        ISS_FAIL(h, ISS_ERR_ARG, "iss_cuda_spectra: bad grid sizes (y - eta_s table: at most 128 points)");
    if (h->nlab <= 0 || !h->d_lab) ISS_FAIL(h, ISS_ERR_STATE, "iss_cuda_spectra: no lab-frame surface uploaded");
    if (opt->include_deltaf_diffusion == 1 && !h->d_kappa)
        ISS_FAIL(h, ISS_ERR_STATE, "iss_cuda_spectra: diffusion delta f needs the kappa_B table");
    cudaSetDevice(h->device);
    const int npt = npT*nphi;

    SpectraArgs A;
    A.lab = h->d_lab;
    A.ncell = h->nlab;
    A.opt = *opt;
    A.tabs = CoefTables{};
    A.tabs.kappa = h->d_kappa;
    A.tabs.gk = h->gk;
    A.npT = npT; A.nphi = nphi; A.ny = ny;
    A.ns = nspecies;
    // chunk size: a function of ncell only (deterministic sums whatever the device or rank count)
    int64_t chunk = 2048;
    while ((A.ncell + chunk - 1)/chunk > 256) chunk *= 2;
    A.chunk = chunk;
    A.nchunk = static_cast<int>((A.ncell + chunk - 1)/chunk);

    // host tables: the trigonometric / hyperbolic caches of the reference's constructor
    // (emissionfunction.cpp:223-249), evaluated with the host libm as the reference does
    std::vector<double> tab(npT + 2*nphi + 3*ny);
    double *t_pT = tab.data(), *t_c = t_pT + npT, *t_s = t_c + nphi, *t_ch = t_s + nphi,
           *t_sh = t_ch + ny, *t_w = t_sh + ny;
    for (int i = 0; i < npT; i++) t_pT[i] = pT[i];
    for (int j = 0; j < nphi; j++) { t_c[j] = cos(phi[j]); t_s[j] = sin(phi[j]); }
    for (int k = 0; k < ny; k++) {
        t_ch[k] = cosh(-y_minus_eta[k]);
        t_sh[k] = sinh(-y_minus_eta[k]);
        t_w[k] = y_weight[k];
    }
    std::vector<DeviceSpecies> ds(nspecies);
    for (int i = 0; i < nspecies; i++) {
        DeviceSpecies d{};
        d.mass = species[i].mass;
        d.mass2 = d.mass*d.mass;
        d.pid = species[i].pid;
        d.gspin = static_cast<int16_t>(species[i].gspin);
        d.baryon = static_cast<int16_t>(species[i].baryon);
        d.strange = static_cast<int16_t>(species[i].strange);
        d.charge = static_cast<int16_t>(species[i].charge);
        d.sign = static_cast<int16_t>(species[i].sign);
        ds[i] = d;
    }
    const size_t tab_bytes = sizeof(double)*tab.size();
    const size_t sp_off = (tab_bytes + 255)/256*256;
    ISS_ENSURE(h, h->d_spec_tab, h->spec_tab_bytes, sp_off + sizeof(DeviceSpecies)*nspecies);
    ISS_CUDA_TRY(h, cudaMemcpyAsync(h->d_spec_tab, tab.data(), tab_bytes, cudaMemcpyHostToDevice, h->stream));
    ISS_CUDA_TRY(h, cudaMemcpyAsync(reinterpret_cast<char *>(h->d_spec_tab) + sp_off, ds.data(),
                                    sizeof(DeviceSpecies)*nspecies, cudaMemcpyHostToDevice, h->stream));
    A.pT = h->d_spec_tab; A.cphi = A.pT + npT; A.sphi = A.cphi + nphi; A.ch = A.sphi + nphi;
    A.sh = A.ch + ny; A.wy = A.sh + ny;
    A.species = reinterpret_cast<const DeviceSpecies *>(reinterpret_cast<char *>(h->d_spec_tab) + sp_off);

    ISS_ENSURE(h, h->d_labrec, h->labrec_bytes, sizeof(double)*REC*A.ncell);
    const size_t n_out = static_cast<size_t>(nspecies)*npt;
    ISS_ENSURE(h, h->d_spec_part, h->spec_part_bytes, sizeof(double)*2*n_out*A.nchunk);
    ISS_ENSURE(h, h->d_spec_out, h->spec_out_bytes, sizeof(double)*2*n_out);
    A.rec = h->d_labrec;
    A.part = h->d_spec_part;
    A.out = h->d_spec_out;

    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0, h->stream);
    spectra_cell_kernel<<<static_cast<unsigned>((A.ncell + 127)/128), 128, 0, h->stream>>>(A);
    ISS_LAUNCHED(h);
    const dim3 grid((npt + SPEC_THREADS - 1)/SPEC_THREADS, nspecies, A.nchunk);
    int bulk = 0;
    if (opt->include_deltaf_bulk == 1 && opt->bulk_deltaf_kind >= 1 && opt->bulk_deltaf_kind <= 4)
        bulk = opt->bulk_deltaf_kind;
    switch (bulk) {
    case 1: launch_spectra<1>(A, grid, h->stream, dN_max != nullptr); break;
    case 2: launch_spectra<2>(A, grid, h->stream, dN_max != nullptr); break;
    case 3: launch_spectra<3>(A, grid, h->stream, dN_max != nullptr); break;
    case 4: launch_spectra<4>(A, grid, h->stream, dN_max != nullptr); break;
    default: launch_spectra<0>(A, grid, h->stream, dN_max != nullptr); break;
    }
    ISS_LAUNCHED(h);
    spectra_reduce_kernel<<<static_cast<unsigned>((n_out + 255)/256), 256, 0, h->stream>>>(A);
    ISS_LAUNCHED(h);
    cudaEventRecord(e1, h->stream);
    ISS_CUDA_TRY(h, cudaGetLastError());
    ISS_CUDA_TRY(h, cudaMemcpyAsync(dN, h->d_spec_out, sizeof(double)*n_out, cudaMemcpyDeviceToHost, h->stream));
    if (dN_max)
        ISS_CUDA_TRY(h, cudaMemcpyAsync(dN_max, h->d_spec_out + n_out, sizeof(double)*n_out,
                                        cudaMemcpyDeviceToHost, h->stream));
    ISS_CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    h->spec_ms = ms;
    h->spec_evals = static_cast<double>(A.ncell)*ny*npt*nspecies;
    return ISS_OK;
}

int iss_cuda_spectra_stats(iss_handle *h, double *evaluations, double *kernel_ms) {
    if (!h) return ISS_ERR_ARG;
    if (evaluations) *evaluations = h->spec_evals;
    if (kernel_ms) *kernel_ms = h->spec_ms;
    return ISS_OK;
}

}  // extern "C"
