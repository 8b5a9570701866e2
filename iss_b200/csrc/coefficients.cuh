// coefficients.cuh -- per-cell delta-f coefficient look-ups and tabulated special
// functions, shared by the yield kernel and the sampler kernel.
// Restates FSSW.cpp:1074-1212 (polynomials), :1379-1430 (14-moment), :1433-1543
// (CE / 22-moment NEoS-BQS), :1571-1606 (kappa_B), :1646-1727 (K_n / E_n lerps).
#ifndef ISS_COEFFICIENTS_CUH_
#define ISS_COEFFICIENTS_CUH_

#include "iss_internal.cuh"
#include "special_functions.cuh"

namespace iss {

struct CoefTables {
    const double *bessel;   // [n][3] K1,K2,K3
    const double *expint;   // [n][9] E2..E18
    SfGrid sf;
    const double *ce;       // [ne*nb][5]
    const double *mom22;    // [ne*nb][8]
    int ce_n;               // 200 (used for both strides, as the reference does)
    const double *mom14;    // [3][nT][nmu]
    Grid2D g14;
    const double *kappa;    // [nT][nmu]
    Grid2D gk;
};

// FSSW::get_special_function_K{1,2,3} (FSSW.cpp:1646-1703): lerp inside
// [x_min, x_max - dx], exact function outside.
__device__ __forceinline__ void sf_index(const SfGrid &g, double arg, int &idx, double &frac) {
    idx = static_cast<int>((arg - g.x_min)/g.dx);
    frac = (arg - g.x_min - idx*g.dx)/g.dx;
}

__device__ __forceinline__ bool sf_in_table(const SfGrid &g, double arg) {
    return !(arg < g.x_min || arg > g.x_max_minus_dx);
}

// NEoS-BQS table bilinear interpolation used by both the CE (5 columns) and the
// 22-moment (8 columns) tables (FSSW.cpp:1433-1543).  ncol = 5 or 8; results for
// columns 2..ncol-1 are written to interp[0..ncol-3].
template <int NCOL>
__device__ __forceinline__ void neos_bqs_interp(const double *__restrict__ tb, int n,
                                                double Edec, double nB, double *interp) {
    const double e0 = __ldg(&tb[0]);
    const double de = __ldg(&tb[static_cast<size_t>(n)*NCOL]) - e0;
    int idx_e = static_cast<int>((Edec - e0)/de);
    idx_e = max(0, min(n - 2, idx_e));
    const int Ne1 = idx_e*n;
    const int Ne2 = (idx_e + 1)*n;
    const double e_frac = (Edec - __ldg(&tb[static_cast<size_t>(Ne1)*NCOL]))/de;
    const double dnB1 = __ldg(&tb[static_cast<size_t>(Ne1 + 1)*NCOL + 1]);
    const double dnB2 = __ldg(&tb[static_cast<size_t>(Ne2 + 1)*NCOL + 1]);
    int idx_nB1 = static_cast<int>(nB/dnB1);
    int idx_nB2 = static_cast<int>(nB/dnB2);
    // the reference clamps only from above; a negative index is undefined
    // behaviour there, here it is clamped to 0 for memory safety.
    idx_nB1 = max(0, min(n - 2, idx_nB1));
    idx_nB2 = max(0, min(n - 2, idx_nB2));
    const double *r1 = tb + static_cast<size_t>(Ne1 + idx_nB1)*NCOL;
    const double *r2 = tb + static_cast<size_t>(Ne2 + idx_nB2)*NCOL;
    const double f1 = fmin(1.0, (nB - __ldg(&r1[1]))/dnB1);
    const double f2 = fmin(1.0, (nB - __ldg(&r2[1]))/dnB2);
#pragma unroll
    for (int i = 2; i < NCOL; i++) {
        const double t1 = __ldg(&r1[i])*(1 - f1) + __ldg(&r1[NCOL + i])*f1;
        const double t2 = __ldg(&r2[i])*(1 - f2) + __ldg(&r2[NCOL + i])*f2;
        interp[i - 2] = t1*(1. - e_frac) + t2*e_frac;
    }
}

// FSSW::getCENEOSBQSCoefficients (FSSW.cpp:1433-1487)
__device__ __forceinline__ void coef_ce(const CoefTables &t, double Edec, double nB, double *c) {
    double ip[3];
    neos_bqs_interp<5>(t.ce, t.ce_n, Edec, nB, ip);
    c[0] = 1./ip[1];
    c[1] = 1./3. - ip[0];
    c[2] = ip[2];
}

// FSSW::get22momNEOSBQSCoefficients (FSSW.cpp:1490-1543)
__device__ __forceinline__ void coef_22mom(const CoefTables &t, double Edec, double nB,
                                           double *c) {
    neos_bqs_interp<8>(t.mom22, t.ce_n, Edec, nB, c);
}

__device__ __forceinline__ double bilinear_clamped(const double *__restrict__ tb, int nx, int ny,
                                                   int ix, int iy, double fx, double fy) {
    const int ix1 = max(0, min(nx - 1, ix)), ix2 = max(0, min(nx - 1, ix + 1));
    const int iy1 = max(0, min(ny - 1, iy)), iy2 = max(0, min(ny - 1, iy + 1));
    const double f1 = __ldg(&tb[ix1*ny + iy1]);
    const double f2 = __ldg(&tb[ix1*ny + iy2]);
    const double f3 = __ldg(&tb[ix2*ny + iy2]);
    const double f4 = __ldg(&tb[ix2*ny + iy1]);
    return f1*(1. - fx)*(1. - fy) + f2*(1. - fx)*fy + f3*fx*fy + f4*fx*(1. - fy);
}

// FSSW::getbulkvisCoefficients(Tdec, mu_B) (FSSW.cpp:1379-1430), OSU 14-moment.
__device__ __forceinline__ void coef_14mom(const CoefTables &t, double T, double muB, double *c) {
    const Grid2D &g = t.g14;
    const int idx_T = static_cast<int>((T - g.x0)/g.dx);
    const int idx_mu = static_cast<int>((muB - g.y0)/g.dy);
    const double fx = (T - g.x0)/g.dx - idx_T;
    const double fy = (muB - g.y0)/g.dy - idx_mu;
    const size_t plane = static_cast<size_t>(g.nx)*g.ny;
    const double c0 = bilinear_clamped(t.mom14, g.nx, g.ny, idx_T, idx_mu, fx, fy);
    const double c1 = bilinear_clamped(t.mom14 + plane, g.nx, g.ny, idx_T, idx_mu, fx, fy);
    const double c2 = bilinear_clamped(t.mom14 + 2*plane, g.nx, g.ny, idx_T, idx_mu, fx, fy);
    const double T3 = T*T*T;
    const double T4 = T3*T;
    c[0] = (c0 - c2)/T4;
    c[1] = c1/T3;
    c[2] = (4.*c2 - c0)/T4;
}

// FSSW::getbulkvisCoefficients(Tdec) (FSSW.cpp:1074-1212): only kind 1 feeds the
// FSSW yield/sampler; kinds 0,2,3,4 are no-ops there.
__device__ __forceinline__ void coef_poly_kind1(double T, double *c) {
    const double x = T/HBARC;
    double p[11];
    p[1] = x;
#pragma unroll
    for (int i = 2; i < 11; i++) p[i] = p[i - 1]*x;
    c[0] = (642096.624265727 - 8163329.49562861*p[1] + 47162768.4292073*p[2]
            - 162590040.002683*p[3] + 369637951.096896*p[4] - 578181331.809836*p[5]
            + 629434830.225675*p[6] - 470493661.096657*p[7] + 230936465.421*p[8]
            - 67175218.4629078*p[9] + 8789472.32652964*p[10]);
    c[1] = (1.18171174036192 - 17.6740645873717*p[1] + 136.298469057177*p[2]
            - 635.999435106846*p[3] + 1918.77100633321*p[4] - 3836.32258307711*p[5]
            + 5136.35746882372*p[6] - 4566.22991441914*p[7] + 2593.45375240886*p[8]
            - 853.908199724349*p[9] + 124.260460450113*p[10]);
    c[2] = 0.0;
}

// FSSW::get_deltaf_qmu_coeff (FSSW.cpp:1571-1606): 1e30 outside the grid.
__device__ __forceinline__ double coef_kappa(const CoefTables &t, double T, double muB) {
    const Grid2D &g = t.gk;
    const int idx_T = static_cast<int>((T - g.x0)/g.dx);
    const int idx_mu = static_cast<int>((muB - g.y0)/g.dy);
    const double fx = (T - g.x0)/g.dx - idx_T;
    const double fy = (muB - g.y0)/g.dy - idx_mu;
    if (idx_mu > g.ny - 2 || idx_T > g.nx - 2 || idx_mu < 0 || idx_T < 0) return 1e30;
    const double f1 = __ldg(&t.kappa[idx_T*g.ny + idx_mu]);
    const double f2 = __ldg(&t.kappa[idx_T*g.ny + idx_mu + 1]);
    const double f3 = __ldg(&t.kappa[(idx_T + 1)*g.ny + idx_mu + 1]);
    const double f4 = __ldg(&t.kappa[(idx_T + 1)*g.ny + idx_mu]);
    return f1*(1. - fx)*(1. - fy) + f2*(1. - fx)*fy + f3*fx*fy + f4*fx*(1. - fy);
}

// The per-cell part of the delta-f set-up that the reference recomputes for every
// species (FSSW.cpp:596-637) and for every sample (FSSW.cpp:982-1008).
struct CellCoef {
    double c[6];        // bulkvisCoefficients
    double kappa;       // deltaf_qmu_coeff
};

struct ModeFlags {
    int include_shear, include_bulk, include_diff;
    int kind;           // bulk_deltaf_kind
    int neos;           // 1: CE (kind 21), 0: 22-moment (kind 20), -1 otherwise
};

__device__ __forceinline__ void cell_coefficients(const CoefTables &t, const ModeFlags &f,
                                                  double Edec, double nB, double T, double muB,
                                                  CellCoef &out) {
#pragma unroll
    for (int i = 0; i < 6; i++) out.c[i] = 0.0;
    out.kappa = 1.0;
    if (f.neos == 1) {
        coef_ce(t, Edec, nB, out.c);
    } else if (f.neos == 0) {
        coef_22mom(t, Edec, nB, out.c);
    }
    if (f.include_bulk == 1 && f.neos == -1) {
        if (f.kind == 11) {
            coef_14mom(t, T, muB, out.c);
        } else if (f.kind == 1) {
            coef_poly_kind1(T, out.c);
        }
    }
    if (f.include_diff == 1) out.kappa = coef_kappa(t, T, muB);
}

}  // namespace iss
#endif  // ISS_COEFFICIENTS_CUH_
