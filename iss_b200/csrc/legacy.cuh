// legacy.cuh -- the reference's legacy "conventional" sampler (MC_sampling = 2,
// EmissionFunctionArray::sample_using_dN_dxtdy_4all_particles_conventional,
// src/emissionfunction.cpp:3273-3623) on the device.  Included by sampler.cu (it shares the cell
// search, the work-item bookkeeping and the Philox protocol of the FSSW path).
//
//   legacy_coef_kernel    per cell: bulk coefficients (getbulkvisCoefficients(T), :3625-3762, kinds
//                         1..4) and kappa_hat (get_deltaf_qmu_coeff, :3788-3823)
//   legacy_yields_kernel  calculate_dN_dxtdy_for_one_particle_species + calculate_dN_analytic
//                         (:2977-3208): Milne-component sigma.u, 10-term series iff m < 0.7,
//                         bulk term only for kind 1, NOT clamped
//   legacy_max_kernel     estimate_maximum (:4006-4153, 4309-4421) for every (species, cell)
//                         (test instrumentation)
//   legacy_sample_kernel  per hadron: cell (RandomVariable1DArray::rand), maximum, then tries of
//                         sample_momemtum_from_a_fluid_cell (:4188-4306) until accepted, emit as
//                         add_one_sampled_particle (:4423-4475).  Persistent lanes: a lane whose
//                         hadron is accepted takes the next work item at once (the acceptance is
//                         ~1/500 per try, so without the refill a warp would idle on its slowest lane).
//                         The lane's cell record lives in shared memory (33 floats per lane), the
//                         per-hadron divisions are done once at the hand-over.  Local charge
//                         conservation: a positive hadron is followed by its conjugate from the same
//                         cell under the same maximum (:3517-3546).
// Random numbers: stream (seed; SAMPLE, species, event, draw), one Philox block per decision:
//   block 0: cell (w0,w1 -> 53 bit) | one block per try: w0 -> pT^2, w1 -> phi, w2 -> y - eta_s,
//   w3 -> accept | after 4999 rejected tries: new cell | boost-invariant: rapidity (w0) |
//   charge-conservation partner: one block per try, never a new cell.
#ifndef ISS_LEGACY_CUH_
#define ISS_LEGACY_CUH_

#include "coefficients.cuh"
#include "legacy_bulk.cuh"

namespace iss {

struct LegacyArgs;
int legacy_args(iss_handle *h, LegacyArgs &G);   // yields.cu

constexpr int LEGACY_THREADS = 256;
constexpr int LEGACY_LAMBERT_N = 40001;     // (200 - 0)/0.005 + 1 (:3858-3870)
constexpr int LEGACY_MAX_IMPATIENCE = 5000;
constexpr long long LEGACY_MAX_TRIES = 20000000;   // safety valve (the reference would never end)

struct LegacyArgs {
    const float *lab;       // [ncell][ISS_LAB_NFIELD]
    const float4 *pos;      // [ncell] x, y, eta_s, 0
    double4 *coef;          // [ncell] c0, c1, c2, kappa_hat
    const double *bulk0;    // [4][nbulk0] T [1/fm], B0, D0, E0 (bulk_deltaf_kind 0), or null
    int nbulk0;
    const double *zx, *zy;  // iSS_tables/z_exp_m_z.dat
    int nz;
    const double *lambert;  // [LEGACY_LAMBERT_N] W0 on x = 0.005 i
    int include_shear, include_bulk, bulk_kind, include_diff, restrict_deltaf;
    double deltaf_max_ratio, pT_to, y_range;
    int64_t ncell, ncell_pad;
    CoefTables tab;         // special-function tables + kappa grid
    double *yields;         // [ns][ncell_pad]
    double *max_out;        // [ns][ncell] (legacy_max_kernel)
};

// principal branch of the Lambert W function, x >= 0 (the reference calls gsl_sf_lambert_W0)
__host__ __device__ inline double legacy_lambert_w0(double x) {
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

static __global__ void legacy_lambert_kernel(double *tab) {
    const int i = blockIdx.x*blockDim.x + threadIdx.x;
    if (i < LEGACY_LAMBERT_N) tab[i] = legacy_lambert_w0(0.005*i);
}

// get_special_function_lambertW (:3936-3953)
__device__ __forceinline__ double legacy_lambertW(const LegacyArgs &A, double arg) {
    const double x_min = 0., x_max = 200.0, dx = 0.005;
    if (arg < x_min || arg > x_max - dx) return legacy_lambert_w0(arg);
    const int idx = static_cast<int>((arg - x_min)/dx);
    const double fraction = (arg - x_min - idx*dx)/dx;
    return (1. - fraction)*__ldg(&A.lambert[idx]) + fraction*__ldg(&A.lambert[idx + 1]);
}

// interpCubicDirect without extrapolation (src/arsenal.cpp:58-110) on an equally spaced table;
// NaN where the reference exit(1)s.  Used as TableFunction::map with interpolation_model 5
// (z_exp_m_z) and as Table::interp(1, col, x, 5) (bulk coefficients of kind 0).
__device__ __forceinline__ double legacy_cubic_direct(const double *__restrict__ x,
                                                      const double *__restrict__ y, int size, double xx) {
    const double x0 = __ldg(&x[0]);
    const double dx = __ldg(&x[1]) - x0;
    if (fabs(xx - x0) < dx*1e-30) return __ldg(&y[0]);
    const long idx = static_cast<long>(floor((xx - x0)/dx));
    if (idx < 0 || idx >= size - 1) return nan("");
    if (idx == 0) {
        const double A0 = __ldg(&y[0]), A1 = __ldg(&y[1]), A2 = __ldg(&y[2]), d = xx - x0;
        return (A0 - 2.0*A1 + A2)/(2.0*dx*dx)*d*d - (3.0*A0 - 4.0*A1 + A2)/(2.0*dx)*d + A0;
    } else if (idx == size - 2) {
        const double A0 = __ldg(&y[size - 3]), A1 = __ldg(&y[size - 2]), A2 = __ldg(&y[size - 1]);
        const double d = xx - (x0 + (idx - 1)*dx);
        return (A0 - 2.0*A1 + A2)/(2.0*dx*dx)*d*d - (3.0*A0 - 4.0*A1 + A2)/(2.0*dx)*d + A0;
    }
    const double A0 = __ldg(&y[idx - 1]), A1 = __ldg(&y[idx]), A2 = __ldg(&y[idx + 1]),
                 A3 = __ldg(&y[idx + 2]);
    const double d = xx - (x0 + idx*dx);
    return (-A0 + 3.0*A1 - 3.0*A2 + A3)/(6.0*dx*dx*dx)*d*d*d + (A0 - 2.0*A1 + A2)/(2.0*dx*dx)*d*d
           - (2.0*A0 + 3.0*A1 - 6.0*A2 + A3)/(6.0*dx)*d + A1;
}

__device__ __forceinline__ double legacy_z_map(const LegacyArgs &A, double xx) {
    return legacy_cubic_direct(A.zx, A.zy, A.nz, xx);
}

// max over E >= mass of E^A f0(E): shared by estimate_ideal_maximum (A = 1, :4006-4053),
// estimate_shear_viscous_maximum (A = 3, :4055-4102), estimate_diffusion_maximum (A = 2, :4104-4153)
static __device__ __noinline__ double legacy_power_max(const LegacyArgs &A, int a, int sign, double mass,
                                                double T, double mu, double f0_mass) {
    const double inv_T = 1./T;
    double massA = mass;
    for (int i = 1; i < a; i++) massA *= mass;
    if (sign == 1) {
        double Emax = T*(legacy_lambertW(A, a*exp(inv_T*mu - a)) + a);
        if (Emax < mass) Emax = mass;
        double EA = Emax;
        for (int i = 1; i < a; i++) EA *= Emax;
        return EA/(exp((Emax - mu)*inv_T) + sign);
    }
    const double rhs = a*exp(inv_T*mu - a);
    if (rhs > 0.3678794) return massA*f0_mass;
    const double Emax = T*(a - legacy_z_map(A, rhs));
    if (Emax < mass) return massA*f0_mass;
    double EA = Emax;
    for (int i = 1; i < a; i++) EA *= Emax;
    const double g1 = EA/(exp((Emax - mu)*inv_T) + sign), g2 = massA*f0_mass;
    return g1 > g2 ? g1 : g2;
}

__device__ __forceinline__ void legacy_load_cell(const LegacyArgs &A, int64_t cell, float *f) {
    const float4 *cr = reinterpret_cast<const float4 *>(A.lab + cell*ISS_LAB_NFIELD);
#pragma unroll
    for (int q = 0; q < ISS_LAB_NFIELD/4; q++) {
        const float4 v = __ldg(cr + q);
        f[4*q] = v.x; f[4*q + 1] = v.y; f[4*q + 2] = v.z; f[4*q + 3] = v.w;
    }
}

// species chemical potential: int x float products summed in float (emissionfunction.cpp:4200, 4327)
__device__ __forceinline__ double legacy_mu(const float *f, int B, int S, int Q) {
    return static_cast<double>(__fadd_rn(
        __fadd_rn(__fmul_rn(static_cast<float>(B), f[ISS_L_MUB]),
                  __fmul_rn(static_cast<float>(S), f[ISS_L_MUS])),
        __fmul_rn(static_cast<float>(Q), f[ISS_L_MUQ])));
}

// EmissionFunctionArray::estimate_maximum (:4309-4421)
static __device__ __noinline__ double legacy_estimate_maximum(const LegacyArgs &A, const float *f,
                                                       const double4 cf, double mass, int sign,
                                                       int degen, int B, int S, int Q) {
    const double prefactor = 1.0/(8.0*(M_PI*M_PI*M_PI))/HBARC/HBARC/HBARC;
    const double Tdec = f[ISS_L_T], inv_Tdec = 1.0/Tdec, Pdec = f[ISS_L_P], Edec = f[ISS_L_E];
    const double mu = legacy_mu(f, B, S, Q);
    double bulkPi = 0.0;
    if (A.include_bulk == 1)
        bulkPi = (A.bulk_kind == 0) ? static_cast<double>(f[ISS_L_BULKPI])
                                    : static_cast<double>(f[ISS_L_BULKPI])/HBARC;
    double prefactor_qmu = 0.0;
    if (A.include_diff == 1) prefactor_qmu = static_cast<double>(f[ISS_L_BN])/(Edec + Pdec);
    // float arithmetic, as the reference's expressions over float members
    const float tau = f[ISS_L_TAU];
    const float uds = __fmul_rn(tau, __fadd_rn(
        __fadd_rn(__fadd_rn(__fmul_rn(f[ISS_L_U0], f[ISS_L_DA0]), __fmul_rn(f[ISS_L_U1], f[ISS_L_DA1])),
                  __fmul_rn(f[ISS_L_U2], f[ISS_L_DA2])),
        __fdiv_rn(__fmul_rn(f[ISS_L_U3], f[ISS_L_DA3]), tau)));
    const float tau2 = __fmul_rn(tau, tau);
    const float dsq = __fmul_rn(tau2, __fsub_rn(
        __fsub_rn(__fsub_rn(__fmul_rn(f[ISS_L_DA0], f[ISS_L_DA0]), __fmul_rn(f[ISS_L_DA1], f[ISS_L_DA1])),
                  __fmul_rn(f[ISS_L_DA2], f[ISS_L_DA2])),
        __fdiv_rn(__fmul_rn(f[ISS_L_DA3], f[ISS_L_DA3]), tau2)));
    const double u_dot_dsigma = uds, dsigma_sq = dsq;
    const double dsigmaT = sqrt(fabs(dsigma_sq - u_dot_dsigma*u_dot_dsigma));
    const double dsigma_all = fabs(u_dot_dsigma) + dsigmaT;
    const double f0_mass = 1./(exp((mass - mu)*inv_Tdec) + sign);
    const double guess_ideal = legacy_power_max(A, 1, sign, mass, Tdec, mu, f0_mass);
    double guess_viscous = 0.0;
    if (A.include_shear == 1) {
        const double p00 = f[ISS_L_PI00], p01 = f[ISS_L_PI01], p02 = f[ISS_L_PI02], p03 = f[ISS_L_PI03],
                     p11 = f[ISS_L_PI11], p12 = f[ISS_L_PI12], p13 = f[ISS_L_PI13], p22 = f[ISS_L_PI22],
                     p23 = f[ISS_L_PI23], p33 = f[ISS_L_PI33];
        const double trace_Pi2 = (p00*p00 + p11*p11 + p22*p22 + p33*p33 - 2.*p01*p01 - 2.*p02*p02
                                  - 2.*p03*p03 + 2.*p12*p12 + 2.*p13*p13 + 2.*p23*p23);
        const double pi_size = sqrt(trace_Pi2)/(Edec + Pdec);
        const double tmp_factor = (sign == -1) ? 2.0 : 1.0;
        guess_viscous = legacy_power_max(A, 3, sign, mass, Tdec, mu, f0_mass)
                        *(tmp_factor/(2.0*Tdec*Tdec)*pi_size);
    }
    double guess_bulk = 0.0;
    if (A.include_bulk == 1)
        guess_bulk = fabs(bulkPi*cf.x)*mass*mass*inv_Tdec/3.*f0_mass*(1. - sign*f0_mass);
    double guess_qmu = 0.0;
    if (A.include_diff == 1) {
        // Vec4 is float (data_struct.h:11)
        const float q2 = __fsub_rn(
            __fsub_rn(__fsub_rn(__fmul_rn(f[ISS_L_Q0], f[ISS_L_Q0]), __fmul_rn(f[ISS_L_Q1], f[ISS_L_Q1])),
                      __fmul_rn(f[ISS_L_Q2], f[ISS_L_Q2])),
            __fmul_rn(f[ISS_L_Q3], f[ISS_L_Q3]));
        const double q_size = sqrt(fabs(static_cast<double>(q2)))/cf.w;
        guess_qmu = prefactor_qmu*legacy_power_max(A, 2, sign, mass, Tdec, mu, f0_mass);
        if (B > 0) guess_qmu += B*guess_ideal;
        guess_qmu *= ((sign == -1) ? 2.0 : 1.0)*q_size;
    }
    return prefactor*degen*dsigma_all*(guess_ideal + guess_viscous + guess_bulk + guess_qmu);
}

static __global__ void legacy_coef_kernel(const LegacyArgs A) {
    const int64_t cell = static_cast<int64_t>(blockIdx.x)*blockDim.x + threadIdx.x;
    if (cell >= A.ncell) return;
    const float *f = A.lab + cell*ISS_LAB_NFIELD;
    const double T = __ldg(&f[ISS_L_T]);
    double c0 = 0., c1 = 0., c2 = 0.;
    if (A.include_bulk == 1 && A.bulk_kind >= 1 && A.bulk_kind <= 4)
        legacy_bulk_poly(A.bulk_kind, T, c0, c1);
    if (A.include_bulk == 1 && A.bulk_kind == 0) {
        // 14-moment coefficients from BulkDf_Coefficients_Hadrons_s95p-v0-PCE.dat
        // (emissionfunction.cpp:3633-3647): B0, E0 in fm^3/GeV^3, D0 in fm^3/GeV^2
        const double T_fm = T/HBARC;
        const int n = A.nbulk0;
        c0 = legacy_cubic_direct(A.bulk0, A.bulk0 + n, n, T_fm)/(HBARC*HBARC*HBARC);
        c1 = legacy_cubic_direct(A.bulk0, A.bulk0 + 2*n, n, T_fm)/(HBARC*HBARC);
        c2 = legacy_cubic_direct(A.bulk0, A.bulk0 + 3*n, n, T_fm)/(HBARC*HBARC*HBARC);
    }
    double kappa = 1.0;
    if (A.include_diff == 1) kappa = coef_kappa(A.tab, T, static_cast<double>(__ldg(&f[ISS_L_MUB])));
    A.coef[cell] = make_double4(c0, c1, c2, kappa);
}

// K_1, K_2 look-ups: lerp inside [x_min, x_max - dx], exact outside (:3875-3911)
__device__ __forceinline__ void legacy_K12(const LegacyArgs &A, double arg, double &K1, double &K2) {
    if (sf_in_table(A.tab.sf, arg)) {
        int idx;
        double fr;
        sf_index(A.tab.sf, arg, idx, fr);
        const double *r0 = A.tab.bessel + 3*static_cast<int64_t>(idx);
        K1 = (1. - fr)*__ldg(&r0[0]) + fr*__ldg(&r0[3]);
        K2 = (1. - fr)*__ldg(&r0[1]) + fr*__ldg(&r0[4]);
    } else {
        double K3;
        bessel_k123(arg, K1, K2, K3);
    }
}

__device__ __forceinline__ double legacy_En(const LegacyArgs &A, double arg, int k) {   // E_{2k+2}
    if (sf_in_table(A.tab.sf, arg)) {
        int idx;
        double fr;
        sf_index(A.tab.sf, arg, idx, fr);
        const double *r0 = A.tab.expint + 9*static_cast<int64_t>(idx);
        return (1. - fr)*__ldg(&r0[k]) + fr*__ldg(&r0[9 + k]);
    }
    return expint_en(2*k + 2, arg);
}

// one thread per cell, loop over the species
static __global__ void __launch_bounds__(128)
legacy_yields_kernel(const LegacyArgs A, const DeviceSpecies *__restrict__ species, int ns) {
    const int64_t cell = static_cast<int64_t>(blockIdx.x)*blockDim.x + threadIdx.x;
    if (cell >= A.ncell) return;
    float f[ISS_LAB_NFIELD];
    legacy_load_cell(A, cell, f);
    const double4 cf = A.coef[cell];
    const double T = f[ISS_L_T], tau = f[ISS_L_TAU], beta = 1./T;
    const double da0 = f[ISS_L_DA0], da1 = f[ISS_L_DA1], da2 = f[ISS_L_DA2], da3 = f[ISS_L_DA3];
    const double sdu = tau*(da0*f[ISS_L_U0] + f[ISS_L_U1]*da1 + f[ISS_L_U2]*da2 + f[ISS_L_U3]*da3/tau);
    double bulkPi = 0.;
    if (A.include_bulk == 1)
        bulkPi = (A.bulk_kind == 0) ? static_cast<double>(f[ISS_L_BULKPI])
                                    : static_cast<double>(f[ISS_L_BULKPI])/HBARC;
    double sdq = 0., pref_q = 0.;
    if (A.include_diff == 1) {
        sdq = tau*(da0*f[ISS_L_Q0] + da1*f[ISS_L_Q1] + da2*f[ISS_L_Q2] + da3*f[ISS_L_Q3]/tau);
        pref_q = static_cast<double>(f[ISS_L_BN])
                 /(static_cast<double>(f[ISS_L_E]) + static_cast<double>(f[ISS_L_P]));
    }
    const double unit_factor = 1.0/(HBARC*HBARC*HBARC);
    // I_1 weights of E_2, E_4, ..., E_18 (:3155-3177): 3/8, then 3 (2k-5)!!/(2^k k!), k = 3..10
    double w[9];
    {
        w[0] = 3./8.;
        double double_factorial = 1., factorial = 2., two_k = 4.;
        for (int k = 3; k <= 10; k++) {
            double_factorial *= (2*k - 5);
            factorial *= k;
            two_k *= 2;
            w[k - 2] = 3.*double_factorial/two_k/factorial;
        }
    }
    for (int s = 0; s < ns; s++) {
        const DeviceSpecies p = species[s];
        const double mass = p.mass;
        const int sign = p.sign;
        const double mu = legacy_mu(f, p.baryon, p.strange, p.charge);
        const double lambda = exp(beta*mu);
        const int trunc = (mass < 0.7) ? 10 : 1;
        double N_eq = 0., b1 = 0., b2 = 0., q1 = 0., q2 = 0.;
        double fug = 1., theta = 1.;
        for (int n = 1; n <= trunc; n++) {
            const double arg = n*mass*beta;
            fug *= lambda;                  // pow(lambda, n)
            if (n > 1) theta *= -sign;      // pow(-sign, n-1)
            double K1, K2;
            legacy_K12(A, arg, K1, K2);
            N_eq += theta/n*fug*K2;
            if (A.include_bulk == 1 && A.bulk_kind == 1) {
                b1 += theta*fug*(mass*beta*K1 + 3./n*K2);
                b2 += theta*fug*K1;
            }
            if (A.include_diff == 1) {
                q1 += theta/n*fug*K2;
                double I = exp(-arg)/arg*(2./(arg*arg) + 2./arg - 1./2.);
                for (int k = 0; k < 9; k++) I += w[k]*legacy_En(A, arg, k);
                const double mbeta = mass*beta;
                I = -(mbeta*mbeta*mbeta)*I;
                q2 += n*theta*fug*I;
            }
        }
        N_eq = mass*mass*T*N_eq;
        b1 = mass*mass/beta*b1;
        b2 = mass*mass*mass/3.*b2;
        q1 = mass*mass/(beta*beta)*q1;
        q2 = 1./(3.*beta*beta*beta)*q2;
        const double prefactor = p.gspin/(2.*M_PI*M_PI);
        double total = unit_factor*prefactor*sdu*N_eq;
        if (A.include_bulk == 1)
            total += unit_factor*prefactor*sdu*(-bulkPi*cf.x)*(-cf.y*b1 + b2);
        if (A.include_diff == 1)
            total += unit_factor*prefactor*sdq/cf.w*(-pref_q*q1 - p.baryon*q2);
        A.yields[static_cast<int64_t>(s)*A.ncell_pad + cell] = total;
    }
}

static __global__ void legacy_clamp_kernel(double *y, int64_t n) {
    const int64_t i = static_cast<int64_t>(blockIdx.x)*blockDim.x + threadIdx.x;
    if (i < n) y[i] = fmax(y[i], 0.);      // RandomVariable1DArray.cpp:38-50 takes max(val, 0)
}

static __global__ void legacy_max_kernel(const LegacyArgs A, const DeviceSpecies *__restrict__ species, int ns) {
    const int64_t i = static_cast<int64_t>(blockIdx.x)*blockDim.x + threadIdx.x;
    if (i >= A.ncell*ns) return;
    const int s = static_cast<int>(i/A.ncell);
    const int64_t cell = i - static_cast<int64_t>(s)*A.ncell;
    float f[ISS_LAB_NFIELD];
    legacy_load_cell(A, cell, f);
    const DeviceSpecies p = species[s];
    A.max_out[i] = legacy_estimate_maximum(A, f, A.coef[cell], p.mass, p.sign, p.gspin, p.baryon,
                                           p.strange, p.charge);
}

// get_deltaf_bulk (:4156-4186)
__device__ __forceinline__ double legacy_deltaf_bulk(int kind, double mass, double pdotu,
                                                     double bulkPi, double Tdec, int sign, double f0,
                                                     const double4 c) {
    const double stat = 1. - sign*f0;
    if (kind == 0) return -stat*bulkPi*(c.x*mass*mass + c.y*pdotu + c.z*pdotu*pdotu);
    const double E_over_T = pdotu/Tdec;
    if (kind == 1) {
        const double mass_over_T = mass/Tdec;
        return -1.0*stat*c.x*(mass_over_T*mass_over_T/(3.*E_over_T) - c.y*E_over_T)*bulkPi;
    }
    if (kind == 2) return -1.*stat*bulkPi*(-c.x + c.y*E_over_T);
    if (kind == 3) return -1.*stat*bulkPi/sqrt(E_over_T)*(-c.x + c.y*E_over_T);
    if (kind == 4) return -1.*stat*bulkPi*(c.x - c.y/E_over_T);
    return 0.0;
}

#ifdef ISS_LEGACY_WITH_SAMPLER
struct LegacyLane {
    int64_t out_slot;
    int64_t cell;
    int s;
    uint32_t event, draw, block;
    int tries;
    long long total_tries;
    double maximum;
    int qsign;          // +1 primary, -1 charge-conservation partner (conjugate quantum numbers)
    double eta_s;       // eta_s of the primary, reused by its partner (:3533-3536)
};

constexpr int LEGACY_LANE_STRIDE = ISS_LAB_NFIELD + 1;    // floats per lane (odd: no bank conflicts)

static __global__ void __launch_bounds__(LEGACY_THREADS, 2)
legacy_sample_kernel(const SamplerArgs A, const LegacyArgs G) {
    extern __shared__ unsigned char smem_raw[];
    DeviceSpecies *sp = reinterpret_cast<DeviceSpecies *>(smem_raw);
    int64_t *sp_off = reinterpret_cast<int64_t *>(sp + A.ns);
    // per-lane copy of the lab-frame cell record the tries read (written once per hadron)
    float *f = reinterpret_cast<float *>(sp_off + A.ns + 1) + threadIdx.x*LEGACY_LANE_STRIDE;
    for (int i = threadIdx.x; i < A.ns; i += blockDim.x) sp[i] = A.species[i];
    for (int i = threadIdx.x; i <= A.ns; i += blockDim.x)
        sp_off[i] = A.off_work[static_cast<int64_t>(i)*A.nev];
    __syncthreads();
    const uint32_t key0 = static_cast<uint32_t>(A.seed), key1 = static_cast<uint32_t>(A.seed >> 32);
    const unsigned full = 0xffffffffu;
    const double prefactor = 1.0/(8.0*(M_PI*M_PI*M_PI)*(HBARC*HBARC*HBARC));
    LegacyLane L;
    L.cell = 0; L.s = 0; L.out_slot = 0; L.event = 0; L.draw = 0; L.block = 0; L.tries = 1;
    L.total_tries = 0; L.maximum = 1.; L.qsign = 1; L.eta_s = 0.;
#pragma unroll
    for (int q = 0; q < ISS_LAB_NFIELD; q++) f[q] = 0.f;
    double4 cf = make_double4(0., 0., 0., 1.);
    // per-hadron constants of the tries (same operations as the reference does per try)
    double c_mu = 0., c_inv_T = 0., c_shear = 0., c_prefq = 0., c_bulkPi = 0.;
    bool busy = false, more = true;
    unsigned long long my_tries = 0, my_redraws = 0, my_giveup = 0;

    // (re)draws the cell of the lane's hadron and prepares the per-cell constants
    auto draw_cell = [&]() {
        uint32_t w0, w1, w2, w3;
        philox_block(L.block++, L.draw, L.event, sample_stream_word3(L.s), key0, key1, w0, w1, w2, w3);
        L.cell = pick_cell(A, L.s, u53(w0, w1));
        legacy_load_cell(G, L.cell, f);
        cf = G.coef[L.cell];
        const DeviceSpecies p = sp[L.s];
        L.maximum = legacy_estimate_maximum(G, f, cf, p.mass, p.sign, p.gspin, p.baryon, p.strange,
                                            p.charge);
        L.tries = 1;
        const double Tdec = f[ISS_L_T];
        const float e_plus_p = __fadd_rn(f[ISS_L_E], f[ISS_L_P]);       // float sum (:4204-4205)
        c_mu = legacy_mu(f, p.baryon, p.strange, p.charge);
        c_inv_T = 1./Tdec;
        c_shear = 1.0/(2.0*Tdec*Tdec*e_plus_p);
        c_prefq = __fdiv_rn(f[ISS_L_BN], e_plus_p);
        c_bulkPi = static_cast<double>(f[ISS_L_BULKPI])/HBARC;
    };

    for (;;) {
        if (!busy && more) {
            const int64_t w = static_cast<int64_t>(atomicAdd(&A.counters[0], 1ull));
            if (w < A.nwork) {
                int s;
                int64_t ev, k;
                work_identity(A, sp_off, w, s, ev, k);
                L.s = s;
                L.event = static_cast<uint32_t>(A.ev_begin + ev);
                L.draw = static_cast<uint32_t>(k);
                L.block = 0u;
                L.total_tries = 0;
                // local charge conservation: a positive hadron is followed by its partner
                const int mult = (A.lcc == 1 && sp[s].charge > 0) ? 2 : 1;
                L.out_slot = __ldg(&A.off_out[ev*A.ns + s]) + k*mult;
                L.qsign = 1;
                draw_cell();
                busy = true;
            } else {
                more = false;
            }
        }
        if (!__any_sync(full, busy || more)) break;
        if (!busy) continue;

        // ---- one try (sample_momemtum_from_a_fluid_cell, :4212-4296)
        const DeviceSpecies p = sp[L.s];
        const double mass = p.mass;
        const int sign = p.sign;
        uint32_t w0, w1, w2, w3;
        philox_block(L.block++, L.draw, L.event, sample_stream_word3(L.s), key0, key1, w0, w1, w2, w3);
        my_tries++;
        L.total_tries++;
        const double Tdec = f[ISS_L_T], inv_Tdec = c_inv_T;
        const double mu = c_mu;
        const double pT = sqrt(G.pT_to*G.pT_to*u32(w0));
        const double u_phi = u32(w1);
        const double yme = (1. - 2.*u32(w2))*G.y_range;
        const double mT = sqrt(mass*mass + pT*pT);
        double sphi, cphi;
        sincospi(2.*u_phi, &sphi, &cphi);
        const double px = pT*cphi, py = pT*sphi;
        const double p0 = mT*cosh(yme), p3 = mT*sinh(yme);
        const double pdotu = p0*f[ISS_L_U0] - px*f[ISS_L_U1] - py*f[ISS_L_U2] - p3*f[ISS_L_U3];
        const double f0 = 1./(exp((pdotu - mu)*inv_Tdec) + sign);
        const double pdsigma = p0*f[ISS_L_DA0] + px*f[ISS_L_DA1] + py*f[ISS_L_DA2]
                               + p3*f[ISS_L_DA3]/f[ISS_L_TAU];
        double delta_f = 0.;
        if (G.include_shear == 1) {
            const double Wfactor = (p0*p0*f[ISS_L_PI00] - 2.0*p0*px*f[ISS_L_PI01] - 2.0*p0*py*f[ISS_L_PI02]
                                    - 2.0*p0*p3*f[ISS_L_PI03] + px*px*f[ISS_L_PI11]
                                    + 2.0*px*py*f[ISS_L_PI12] + 2.0*px*p3*f[ISS_L_PI13]
                                    + py*py*f[ISS_L_PI22] + 2.0*py*p3*f[ISS_L_PI23] + p3*p3*f[ISS_L_PI33]);
            delta_f += (1. - sign*f0)*Wfactor*c_shear;
        }
        if (G.include_bulk == 1)
            delta_f += legacy_deltaf_bulk(G.bulk_kind, mass, pdotu, c_bulkPi, Tdec, sign, f0, cf);
        if (G.include_diff == 1) {
            const double qmufactor = p0*f[ISS_L_Q0] - px*f[ISS_L_Q1] - py*f[ISS_L_Q2] - p3*f[ISS_L_Q3];
            delta_f += (1. - sign*f0)*(c_prefq - (L.qsign*p.baryon)/pdotu)*qmufactor/cf.w;
        }
        double resize_factor = 1.0;
        if (G.restrict_deltaf == 1)
            resize_factor = fmin(1., G.deltaf_max_ratio/(fabs(delta_f) + 1e-10));
        const double result = prefactor*p.gspin*f0*pdsigma*f[ISS_L_TAU]*(1. + delta_f*resize_factor);
        const double accept_prob = result/L.maximum;
        if (u32(w3) < accept_prob) {
            // ---- accepted: add_one_sampled_particle (:4423-4475)
            const float4 pos = __ldg(&G.pos[L.cell]);
            if (L.qsign > 0) {
                double eta_s = pos.z;
                if (A.hydro_mode != 2) {
                    uint32_t r0, r1, r2, r3;
                    philox_block(L.block++, L.draw, L.event, sample_stream_word3(L.s), key0, key1,
                                 r0, r1, r2, r3);
                    const double rap = A.y_LB + (A.y_RB - A.y_LB)*u32(r0);
                    eta_s = rap - yme;
                }
                L.eta_s = eta_s;
            }
            const double eta_s = L.eta_s;
            const double rapidity_y = yme + eta_s;
            const double tau = f[ISS_L_TAU];
            float2 *dst = reinterpret_cast<float2 *>(A.out + L.out_slot);
            dst[0] = make_float2(__int_as_float(L.qsign > 0 ? p.pid : -p.pid), static_cast<float>(mass));
            dst[1] = make_float2(static_cast<float>(mT*cosh(rapidity_y)), static_cast<float>(px));
            dst[2] = make_float2(static_cast<float>(py), static_cast<float>(mT*sinh(rapidity_y)));
            dst[3] = make_float2(static_cast<float>(tau*cosh(eta_s)), pos.x);
            dst[4] = make_float2(pos.y, static_cast<float>(tau*sinh(eta_s)));
            if (A.trace_cell) {
                A.trace_cell[L.out_slot] = static_cast<int32_t>(L.cell);
                A.trace_tries[L.out_slot] = static_cast<int32_t>(L.total_tries);
            }
            busy = false;
            if (A.lcc == 1 && L.qsign > 0 && p.charge > 0) {
                // a negative partner from the SAME cell with the SAME maximum (:3517-3546)
                L.qsign = -1;
                L.out_slot += 1;
                L.tries = 1;
                L.total_tries = 0;
                c_mu = legacy_mu(f, -p.baryon, -p.strange, -p.charge);
                busy = true;
            }
        } else {
            L.tries++;
            if (L.tries >= LEGACY_MAX_IMPATIENCE) {
                if (L.total_tries > LEGACY_MAX_TRIES) {
                    float2 *dst = reinterpret_cast<float2 *>(A.out + L.out_slot);
#pragma unroll
                    for (int q = 0; q < 5; q++) dst[q] = make_float2(0.f, 0.f);
                    my_giveup++;
                    busy = false;
                } else if (L.qsign > 0) {
                    // status 0 -> `continue`: a NEW cell is drawn (:3478-3480)
                    my_redraws++;
                    draw_cell();
                } else {
                    L.tries = 1;    // the partner's do-while stays in the cell (:3521-3528)
                }
            }
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        my_tries += __shfl_down_sync(full, my_tries, d);
        my_redraws += __shfl_down_sync(full, my_redraws, d);
        my_giveup += __shfl_down_sync(full, my_giveup, d);
    }
    if ((threadIdx.x & 31) == 0) {
        if (my_giveup) atomicAdd(&A.counters[6], my_giveup);
        if (my_tries) atomicAdd(&A.counters[1], my_tries);
        if (my_redraws) atomicAdd(&A.counters[2], my_redraws);
    }
}

#endif  // ISS_LEGACY_WITH_SAMPLER

}  // namespace iss
#endif  // ISS_LEGACY_CUH_
