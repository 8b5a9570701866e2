// ingest.cu -- freeze-out surface ingest on the device.
//
// Replaces, for binary MUSIC surfaces, the per-cell work of
//   read_FOdata::read_FOsurfdat_MUSIC / _boost_invariant   (readindata.cpp:395-546, 626-765)
//   read_FOdata::regulate_surface_cells + getValuesFromHRGEOS + regulate_Wmunu
//                                                           (readindata.cpp:768-842, 1216-1309)
//   iSS::computeFOSurfTmunu (per-cell tensors)              (iSS.cpp:378-445)
//   iSS::transform_to_local_rest_frame                      (iSS.cpp:170-293)
// One thread per file record: parse -> T filter -> regulation -> T^{mu nu} tensor -> LRF record
// -> u.dsigma filter; two order-preserving compactions follow.  The arithmetic restates the host
// path (iss_b200/host/readindata.cpp, iSS.cpp) expression by expression, including the mixed
// float / double types of the reference; this file is compiled with -fmad=false so that no
// product-sum is contracted.  Only cosh/sinh of the space-time rapidity come from a different
// math library than the host's (both are rounded to float: a difference needs the double result
// to fall within ~2 ulp of a float rounding boundary, ~1e-8 per value).
#include "iss_internal.cuh"

#include <cmath>

namespace iss {

namespace {

constexpr int RAW = 34;         // floats per file record
constexpr int TM = 16;

struct LabCell {
    float tau, xpt, ypt, eta;
    float da0, da1, da2, da3;
    float u0, u1, u2, u3;
    float Edec, Tdec, Pdec;
    float Bn, muB, muS, muQ;
    float pi00, pi01, pi02, pi03, pi11, pi12, pi13, pi22, pi23, pi33;
    float bulkPi;
    float qmu0, qmu1, qmu2, qmu3;
};

struct IngestArgs {
    const float *raw;           // [n][34]
    int64_t n;
    iss_ingest_options opt;
    const double *hrg;          // [rows][7]: ed, nB, P, T, muB, muS, muQ
    float *lrf_tmp;             // [n][28]
    float *tm_tmp;              // [n][16]
    int64_t *keepT, *keep;      // flags, later exclusive prefixes
    uint8_t *status;            // [n]
    float *lrf_out, *tm_out;    // compacted
};

// Milne -> (t, z) rotation; float members as in the reference (iSS.cpp:189-190)
struct EtaRot {
    float ch, sh;
    __device__ explicit EtaRot(float eta)
        : ch(static_cast<float>(cosh(static_cast<double>(eta)))),
          sh(static_cast<float>(sinh(static_cast<double>(eta)))) {}
    __device__ float time_like(float a0, float a3) const { return a0*ch + a3*sh; }
    __device__ float z_like(float a0, float a3) const { return a3*ch + a0*sh; }
};

// readindata.cpp:646-689 (binary record, file units -> GeV)
__device__ void parse_cell(const float *a, bool boost_inv, LabCell &s) {
    s.tau = a[0]; s.xpt = a[1]; s.ypt = a[2];
    s.eta = boost_inv ? 0.0f : a[3];
    s.da0 = a[4]; s.da1 = a[5]; s.da2 = a[6];
    s.da3 = boost_inv ? 0.0f : a[7];
    s.u0 = a[8]; s.u1 = a[9]; s.u2 = a[10]; s.u3 = a[11];
    s.Edec = static_cast<float>(a[12]*HBARC);
    s.Tdec = static_cast<float>(a[13]*HBARC);
    s.muB = static_cast<float>(a[14]*HBARC);
    s.muS = static_cast<float>(a[15]*HBARC);
    s.muQ = static_cast<float>(a[16]*HBARC);
    s.Pdec = a[17]*s.Tdec - s.Edec;             // float arithmetic (readindata.cpp:670)
    s.pi00 = static_cast<float>(a[18]*HBARC); s.pi01 = static_cast<float>(a[19]*HBARC);
    s.pi02 = static_cast<float>(a[20]*HBARC); s.pi03 = static_cast<float>(a[21]*HBARC);
    s.pi11 = static_cast<float>(a[22]*HBARC); s.pi12 = static_cast<float>(a[23]*HBARC);
    s.pi13 = static_cast<float>(a[24]*HBARC); s.pi22 = static_cast<float>(a[25]*HBARC);
    s.pi23 = static_cast<float>(a[26]*HBARC); s.pi33 = static_cast<float>(a[27]*HBARC);
    s.bulkPi = static_cast<float>(a[28]*HBARC);
    s.Bn = a[29];
    s.qmu0 = a[30]; s.qmu1 = a[31]; s.qmu2 = a[32]; s.qmu3 = a[33];
}

// bilinear (e, n_B) interpolation of the HRG table (readindata.cpp:1249-1309); eos = {P, T, muB,
// muS, muQ}; -1: energy density outside the table
__device__ int hrg_lookup(const double *hrg, long rows, int nBlen, double ed, double nB, double eos[5]) {
    for (int i = 0; i < 5; i++) eos[i] = 0.;
    auto H = [&](long row, int col) { return hrg[row*7 + col]; };
    const double de = H(nBlen, 0) - H(0, 0);
    const double e0 = H(0, 0);
    const int e_idx = static_cast<int>((ed - e0)/de);
    if (e_idx < 0 || e_idx >= static_cast<int>(rows/nBlen) - 2) return -1;
    const long r1 = static_cast<long>(e_idx)*nBlen;
    const long r2 = static_cast<long>(e_idx + 1)*nBlen;
    const double e_frac = (ed - H(r1, 0))/de;
    double f1 = 0, f2 = 0;
    int i1 = 0, i2 = 0;
    if (nBlen > 1) {
        const double dnB1 = H(r1 + 1, 1), dnB2 = H(r2 + 1, 1);
        i1 = min(nBlen - 2, static_cast<int>(nB/dnB1));
        i2 = min(nBlen - 2, static_cast<int>(nB/dnB2));
        f1 = fmin(1., (nB - H(r1 + i1, 1))/dnB1);
        f2 = fmin(1., (nB - H(r2 + i2, 1))/dnB2);
    }
    auto interp = [&](int col) {
        const double a = H(r1 + i1, col)*(1. - f1) + H(r1 + i1 + 1, col)*f1;
        const double b = H(r2 + i2, col)*(1. - f2) + H(r2 + i2 + 1, col)*f2;
        return a*(1 - e_frac) + b*e_frac;
    };
    eos[0] = interp(2);
    eos[1] = interp(3);
    if (nBlen > 1)
        for (int c = 4; c < 7; c++) eos[c - 2] = interp(c);
    return 0;
}

// transverse, traceless projection (readindata.cpp:1216-1246)
__device__ void regulate_Wmunu(const double u[4], const double W[4][4], double R[4][4]) {
    const double g[4] = {-1., 1., 1., 1.};
    double u_dot_pi[4], u_mu[4];
    for (int i = 0; i < 4; i++) {
        u_dot_pi[i] = -u[0]*W[0][i] + u[1]*W[1][i] + u[2]*W[2][i] + u[3]*W[3][i];
        u_mu[i] = g[i]*u[i];
    }
    const double tr_pi = -W[0][0] + W[1][1] + W[2][2] + W[3][3];
    double upu = 0.0;
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) upu += u_mu[i]*W[i][j]*u_mu[j];
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) {
            const double gij = (i == j) ? g[i] : 0.;
            R[i][j] = (W[i][j] + u[i]*u_dot_pi[j] + u[j]*u_dot_pi[i] + u[i]*u[j]*upu
                       - 1./3.*(gij + u[i]*u[j])*(tr_pi + upu));
        }
}

// readindata.cpp:768-842; returns -1 when the EOS table does not cover the cell
__device__ int regulate_cell(LabCell &s, const iss_ingest_options &o, const double *hrg) {
    int rc = 0;
    if (o.regulate_eos) {
        double eos[5];
        rc = hrg_lookup(hrg, static_cast<long>(o.hrg_rows), o.hrg_nB, s.Edec, s.Bn, eos);
        if (rc == 0) {
            s.Tdec = static_cast<float>(eos[1]);
            s.muB = static_cast<float>(eos[2]);
            s.muS = static_cast<float>(eos[3]);
            s.muQ = static_cast<float>(eos[4]);
            s.Pdec = static_cast<float>(eos[0]);
        }
    }
    // 1. is a double literal, the products are float (readindata.cpp:796-798)
    s.u0 = static_cast<float>(sqrt(1. + s.u1*s.u1 + s.u2*s.u2 + s.u3*s.u3));
    s.qmu0 = (s.u1*s.qmu1 + s.u2*s.qmu2 + s.u3*s.qmu3)/s.u0;
    double u[4] = {s.u0, s.u1, s.u2, s.u3};
    double W[4][4] = {{s.pi00, s.pi01, s.pi02, s.pi03},
                      {s.pi01, s.pi11, s.pi12, s.pi13},
                      {s.pi02, s.pi12, s.pi22, s.pi23},
                      {s.pi03, s.pi13, s.pi23, s.pi33}};
    double R[4][4];
    regulate_Wmunu(u, W, R);
    s.pi00 = static_cast<float>(R[0][0]); s.pi01 = static_cast<float>(R[0][1]);
    s.pi02 = static_cast<float>(R[0][2]); s.pi03 = static_cast<float>(R[0][3]);
    s.pi11 = static_cast<float>(R[1][1]); s.pi12 = static_cast<float>(R[1][2]);
    s.pi13 = static_cast<float>(R[1][3]); s.pi22 = static_cast<float>(R[2][2]);
    s.pi23 = static_cast<float>(R[2][3]); s.pi33 = static_cast<float>(R[3][3]);
    return rc;
}

// symmetric pi^{mu nu} in (t,x,y,z) components (iSS.cpp:246-268); the `2.` literals make those
// terms double
__device__ void shear_to_tz(const LabCell &c, const EtaRot &r, float out[4][4]) {
    const float ch = r.ch, sh = r.sh;
    out[0][0] = (c.pi00*ch*ch + 2.*c.pi03*ch*sh + c.pi33*sh*sh);
    out[0][1] = c.pi01*ch + c.pi13*sh;
    out[0][2] = c.pi02*ch + c.pi23*sh;
    out[0][3] = (c.pi00*ch*sh + c.pi03*(ch*ch + sh*sh) + c.pi33*sh*ch);
    out[1][1] = c.pi11;
    out[1][2] = c.pi12;
    out[1][3] = c.pi01*sh + c.pi13*ch;
    out[2][2] = c.pi22;
    out[2][3] = c.pi02*sh + c.pi23*ch;
    out[3][3] = (c.pi00*sh*sh + 2.*c.pi03*sh*ch + c.pi33*ch*ch);
    for (int i = 1; i < 4; i++)
        for (int j = 0; j < i; j++) out[i][j] = out[j][i];
}

// y[i] = sum_j L[i][j] x[j], accumulated in float like the reference's Vec4 += double
__device__ void boost_apply(const double L[4][4], const float x[4], float y[4]) {
    for (int i = 0; i < 4; i++) {
        y[i] = 0.f;
        for (int j = 0; j < 4; j++) y[i] += L[i][j]*x[j];
    }
}

// iSS.cpp:378-445, one cell
__device__ void cell_tmunu(const LabCell &c, const EtaRot &rot, const float pi_tz[4][4], float *out) {
    const float u[4] = {rot.time_like(c.u0, c.u3), c.u1, c.u2, rot.z_like(c.u0, c.u3)};
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) {
            const float gij = (i != j) ? 0.f : (i == 0 ? 1.f : -1.f);
            out[4*i + j] = (c.Edec*u[i]*u[j] - (c.Pdec + c.bulkPi)*(gij - u[i]*u[j]) + pi_tz[i][j]);
        }
}

// iSS.cpp:170-293, one cell; rec in ISS_F_* order; returns false when u.dsigma < 0 (iSS.cpp:226)
__device__ bool lrf_transform(const LabCell &c, const EtaRot &rot, const float pi_tz[4][4], float *rec) {
    const float ut = rot.time_like(c.u0, c.u3);
    const float uz = rot.z_like(c.u0, c.u3);
    const float ux = c.u1, uy = c.u2;
    const double g = ut + 1.;
    const double L[4][4] = {{ut, -ux, -uy, -uz},
                            {-ux, 1. + ux*ux/g, ux*uy/g, ux*uz/g},
                            {-uy, ux*uy/g, 1. + uy*uy/g, uy*uz/g},
                            {-uz, ux*uz/g, uy*uz/g, 1. + uz*uz/g}};
    // contravariant surface normal in (t,x,y,z) from the Milne covariant components
    const float dsigma[4] = {c.tau*c.da0*rot.ch - c.da3*rot.sh, -c.tau*c.da1, -c.tau*c.da2,
                             -c.da3*rot.ch + c.tau*c.da0*rot.sh};
    float ds[4];
    boost_apply(L, dsigma, ds);
    const bool keep = !(ds[0] < 0);
    rec[ISS_F_TAU] = c.tau; rec[ISS_F_X] = c.xpt; rec[ISS_F_Y] = c.ypt; rec[ISS_F_ETA] = c.eta;
    rec[ISS_F_DA0] = ds[0]; rec[ISS_F_DA1] = -ds[1]; rec[ISS_F_DA2] = -ds[2]; rec[ISS_F_DA3] = -ds[3];
    rec[ISS_F_UT] = ut; rec[ISS_F_UX] = ux; rec[ISS_F_UY] = uy; rec[ISS_F_UZ] = uz;
    rec[ISS_F_E] = c.Edec; rec[ISS_F_T] = c.Tdec; rec[ISS_F_P] = c.Pdec; rec[ISS_F_NB] = c.Bn;
    rec[ISS_F_MUB] = c.muB; rec[ISS_F_MUS] = c.muS; rec[ISS_F_MUQ] = c.muQ;
    rec[ISS_F_BULKPI] = c.bulkPi;
    const float q_tz[4] = {rot.time_like(c.qmu0, c.qmu3), c.qmu1, c.qmu2, rot.z_like(c.qmu0, c.qmu3)};
    float q[4];
    boost_apply(L, q_tz, q);
    rec[ISS_F_QX] = q[1]; rec[ISS_F_QY] = q[2]; rec[ISS_F_QZ] = q[3];
    float pi_lrf[4][4];
    for (int i = 1; i < 3; i++)             // only xx, xy, xz, yy, yz are kept
        for (int j = i; j < 4; j++) {
            float acc = 0.;
            for (int a = 0; a < 4; a++)
                for (int b = 0; b < 4; b++) acc += (L[i][a]*pi_tz[a][b]*L[b][j]);
            pi_lrf[i][j] = acc;
        }
    rec[ISS_F_PIXX] = pi_lrf[1][1]; rec[ISS_F_PIXY] = pi_lrf[1][2]; rec[ISS_F_PIXZ] = pi_lrf[1][3];
    rec[ISS_F_PIYY] = pi_lrf[2][2]; rec[ISS_F_PIYZ] = pi_lrf[2][3];
    return keep;
}

__global__ void __launch_bounds__(128)
ingest_kernel(const IngestArgs A) {
    const int64_t cell = static_cast<int64_t>(blockIdx.x)*blockDim.x + threadIdx.x;
    if (cell >= A.n) return;
    float a[RAW];
    const float2 *src = reinterpret_cast<const float2 *>(A.raw + cell*RAW);     // 136-byte records
#pragma unroll
    for (int q = 0; q < RAW/2; q++) {
        const float2 v = __ldg(src + q);
        a[2*q] = v.x;
        a[2*q + 1] = v.y;
    }
    LabCell s;
    parse_cell(a, A.opt.boost_invariant != 0, s);
    uint8_t status = 0;
    const bool keepT = s.Tdec > 0.01;          // readindata.cpp:752
    bool keep = false;
    if (keepT) {
        if (regulate_cell(s, A.opt, A.hrg) != 0) status |= ISS_INGEST_EOS_RANGE;
        const EtaRot rot(s.eta);
        float pi_tz[4][4];
        shear_to_tz(s, rot, pi_tz);
        float tm[TM];
        cell_tmunu(s, rot, pi_tz, tm);
        float4 *td = reinterpret_cast<float4 *>(A.tm_tmp + cell*TM);
#pragma unroll
        for (int q = 0; q < 4; q++) td[q] = make_float4(tm[4*q], tm[4*q + 1], tm[4*q + 2], tm[4*q + 3]);
        float rec[ISS_NFIELD];
        keep = lrf_transform(s, rot, pi_tz, rec);
        if (!keep) status |= ISS_INGEST_DROPPED_NORMAL;
        float4 *rd = reinterpret_cast<float4 *>(A.lrf_tmp + cell*ISS_NFIELD);
#pragma unroll
        for (int q = 0; q < ISS_NFIELD/4; q++)
            rd[q] = make_float4(rec[4*q], rec[4*q + 1], rec[4*q + 2], rec[4*q + 3]);
    } else {
        status |= ISS_INGEST_DROPPED_T;
    }
    A.keepT[cell] = keepT ? 1 : 0;
    A.keep[cell] = keep ? 1 : 0;
    A.status[cell] = status;
}

// order-preserving compaction with the exclusive prefixes of the two flag arrays
__global__ void __launch_bounds__(256)
ingest_compact_kernel(const IngestArgs A, const int64_t *posT, const int64_t *pos) {
    const int64_t i = static_cast<int64_t>(blockIdx.x)*blockDim.x + threadIdx.x;
    const int64_t cell = i >> 3;            // 8 threads per cell: float4 lanes
    const int lane = static_cast<int>(i & 7);
    if (cell >= A.n) return;
    const uint8_t st = A.status[cell];
    if (st & ISS_INGEST_DROPPED_T) return;
    if (lane < 4)
        reinterpret_cast<float4 *>(A.tm_out + posT[cell]*TM)[lane] =
            reinterpret_cast<const float4 *>(A.tm_tmp + cell*TM)[lane];
    if (!(st & ISS_INGEST_DROPPED_NORMAL) && lane < ISS_NFIELD/4)
        reinterpret_cast<float4 *>(A.lrf_out + pos[cell]*ISS_NFIELD)[lane] =
            reinterpret_cast<const float4 *>(A.lrf_tmp + cell*ISS_NFIELD)[lane];
}

}  // namespace

}  // namespace iss

using namespace iss;

extern "C" {

int iss_cuda_ingest_music_binary(iss_handle *h, const float *raw, int64_t ncell,
                                 const iss_ingest_options *opt, const double *hrg, float *lrf_out,
                                 float *tmunu_out, uint8_t *status_out, iss_ingest_result *res) {
    if (!h || !raw || ncell <= 0 || !opt || !lrf_out || !res) return ISS_ERR_ARG;
    if (opt->regulate_eos && (!hrg || opt->hrg_rows <= 0 || opt->hrg_nB <= 0))
        ISS_FAIL(h, ISS_ERR_ARG, "iss_cuda_ingest_music_binary: regulation needs the HRG table");
    cudaSetDevice(h->device);
    const size_t n = static_cast<size_t>(ncell);
    // one arena: raw | lrf_tmp | tm_tmp | lrf_out | tm_out | keepT | keep | posT | pos | status | hrg
    const size_t b_raw = sizeof(float)*RAW*n, b_lrf = sizeof(float)*ISS_NFIELD*n, b_tm = sizeof(float)*TM*n;
    const size_t b_i64 = sizeof(int64_t)*(n + 1);
    const size_t b_hrg = opt->regulate_eos ? sizeof(double)*7*static_cast<size_t>(opt->hrg_rows) : 0;
    auto up = [](size_t v) { return (v + 255)/256*256; };
    const size_t total = up(b_raw) + 2*up(b_lrf) + 2*up(b_tm) + 4*up(b_i64) + up(n) + up(b_hrg);
    ISS_ENSURE(h, h->d_ingest, h->ingest_bytes, total);
    char *p = static_cast<char *>(h->d_ingest);
    auto take = [&](size_t bytes) { char *q = p; p += up(bytes); return q; };
    IngestArgs A;
    float *d_raw = reinterpret_cast<float *>(take(b_raw));
    A.raw = d_raw;
    A.lrf_tmp = reinterpret_cast<float *>(take(b_lrf));
    A.tm_tmp = reinterpret_cast<float *>(take(b_tm));
    A.lrf_out = reinterpret_cast<float *>(take(b_lrf));
    A.tm_out = reinterpret_cast<float *>(take(b_tm));
    A.keepT = reinterpret_cast<int64_t *>(take(b_i64));
    A.keep = reinterpret_cast<int64_t *>(take(b_i64));
    int64_t *d_posT = reinterpret_cast<int64_t *>(take(b_i64));
    int64_t *d_pos = reinterpret_cast<int64_t *>(take(b_i64));
    A.status = reinterpret_cast<uint8_t *>(take(n));
    double *d_hrg = reinterpret_cast<double *>(take(b_hrg));
    A.hrg = d_hrg;
    A.n = ncell;
    A.opt = *opt;
    ISS_CUDA_TRY(h, cudaMemcpyAsync(d_raw, raw, b_raw, cudaMemcpyHostToDevice, h->stream));
    ISS_CUDA_TRY(h, cudaMemsetAsync(A.keepT + ncell, 0, sizeof(int64_t), h->stream));
    ISS_CUDA_TRY(h, cudaMemsetAsync(A.keep + ncell, 0, sizeof(int64_t), h->stream));
    if (b_hrg) ISS_CUDA_TRY(h, cudaMemcpyAsync(d_hrg, hrg, b_hrg, cudaMemcpyHostToDevice, h->stream));
    ingest_kernel<<<static_cast<unsigned>((ncell + 127)/128), 128, 0, h->stream>>>(A); ISS_LAUNCHED(h);
    ISS_CUDA_TRY(h, cudaGetLastError());
    int64_t nT = 0, nkeep = 0;
    int rc = device_exclusive_scan_i64(h, A.keepT, d_posT, ncell, &nT);
    if (rc) return rc;
    rc = device_exclusive_scan_i64(h, A.keep, d_pos, ncell, &nkeep);
    if (rc) return rc;
    ingest_compact_kernel<<<static_cast<unsigned>((ncell*8 + 255)/256), 256, 0, h->stream>>>(A, d_posT, d_pos);
    ISS_LAUNCHED(h);
    ISS_CUDA_TRY(h, cudaGetLastError());
    if (nkeep > 0)
        ISS_CUDA_TRY(h, cudaMemcpyAsync(lrf_out, A.lrf_out, sizeof(float)*ISS_NFIELD*nkeep,
                                        cudaMemcpyDeviceToHost, h->stream));
    if (tmunu_out && nT > 0)
        ISS_CUDA_TRY(h, cudaMemcpyAsync(tmunu_out, A.tm_out, sizeof(float)*TM*nT, cudaMemcpyDeviceToHost,
                                        h->stream));
    if (status_out)
        ISS_CUDA_TRY(h, cudaMemcpyAsync(status_out, A.status, n, cudaMemcpyDeviceToHost, h->stream));
    ISS_CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    res->n_in = ncell;
    res->n_after_T = nT;
    res->n_kept = nkeep;
    return ISS_OK;
}

}  // extern "C"
