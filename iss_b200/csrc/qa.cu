// qa.cu -- QA observables on the device (the only thing NCCL reduces).
//
// Replaces iSS::perform_checks + Histogram (iSS.cpp:59-83, Histogram.cpp:7-61),
// iSS::construct_Tmunu_from_particle_samples (iSS.cpp:296-363) and
// FSSW::computeAvgTotalEnergyMomentum (FSSW.cpp:2028-2059); adds the y, phi, v2 and
// per-species multiplicity moments of SURVEY appendix E.  One CTA per event: the
// event's hadrons are binned in shared memory, then added to the global block with
// one atomicAdd per non-empty bin.  Every entry of the block is a plain sum, so the
// block can be all-reduced across ranks.  (FP64 atomics make the last bits of the
// floating-point sums order dependent; counts are exact.)
#include "iss_internal.cuh"

namespace iss {

struct QaArgs {
    const iss_hadron *hadrons;
    const int64_t *event_off;
    int64_t nev;
    int32_t pids[ISS_QA_NSPEC];
    int npid;
    const iss_decay_species *dsp;
    int ndsp;
    double *qa;
    double *scratch;        // [gridDim.x][QA_SCRATCH] zeroed: per-CTA FP64 sums (pT per bin, v2 numerators)
};

constexpr int QA_THREADS = 256;
constexpr int QA_CTAS_PER_SM = 3;       // 85 registers per thread, 69 KB of shared memory per CTA
constexpr int QA_HASH_BITS = 11, QA_HASH = 1 << QA_HASH_BITS;
// FP64 sums per (tracked species, bin).  Shared memory has no native FP64 add (a compare-and-swap
// loop: a third of the kernel's stall samples when these sums lived there), global memory has one
// that needs no answer (RED.ADD.F64 at the L2): every CTA adds into its own L2-resident block and
// folds it into the QA block at the end.
constexpr int QA_SCRATCH = ISS_QA_NSPEC*(ISS_QA_NPT + ISS_QA_NV2);

__device__ __forceinline__ double block_sum(double v, double *red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_down_sync(0xffffffffu, v, d);
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    double t = 0.;
    if (threadIdx.x == 0)
        for (int w = 0; w < QA_THREADS/32; w++) t += red[w];
    return t;   // valid on thread 0
}

// shared-memory image of the per-species histograms of one CTA (flushed once at the end)
struct QaShared {
    int pt_evt[ISS_QA_NSPEC][ISS_QA_NPT];       // current event only (for the per-event squares)
    int n_evt[ISS_QA_NSPEC];
    unsigned long long pt_cnt[ISS_QA_NSPEC][ISS_QA_NPT];
    unsigned long long pt_sq[ISS_QA_NSPEC][ISS_QA_NPT];
    unsigned long long n_tot[ISS_QA_NSPEC], n_sq[ISS_QA_NSPEC];
    unsigned int y_cnt[ISS_QA_NSPEC][ISS_QA_NY];
    unsigned int phi_cnt[ISS_QA_NSPEC][ISS_QA_NPHI];
    unsigned int v2_den[ISS_QA_NSPEC][ISS_QA_NV2];
    double red[QA_THREADS/32];
    double P1[4], P2[4];        // sum over this CTA's events of P^mu and of its square (thread 0)
    long long n_had, n_evt_done;
    // bin edges of the rapidity and azimuth histograms as the quantities the bins are decided on:
    // sinh(y_i) (compared with p_z/m_T) and the unit vectors of the sector boundaries (sign of a
    // cross product), so that no asinh / atan2 in double precision is needed per hadron
    double y_edge[ISS_QA_NY + 1];
    double phi_cos[ISS_QA_NPHI + 1], phi_sin[ISS_QA_NPHI + 1];
    // pid -> (tracked slot, B, S, Q): open-addressing hash, built once per CTA.  Hadrons are
    // species-ordered inside an event, but most of the ~300 species have fewer hadrons per event
    // than the CTA has threads, so nearly every record a thread reads has a new pid.
    int hash_key[QA_HASH];
    int hash_val[QA_HASH];
};

// (k + 1) | B << 8 | S << 16 | Q << 24 with signed 8-bit charges; k = -1: not tracked
__device__ __forceinline__ int qa_pack(int k, int b, int sq, int q) {
    return ((k + 1) & 0xff) | ((b & 0xff) << 8) | ((sq & 0xff) << 16) | ((q & 0xff) << 24);
}
__device__ __forceinline__ unsigned qa_hash(int pid) {
    return (static_cast<unsigned>(pid)*2654435761u) >> (32 - QA_HASH_BITS);
}
__device__ __forceinline__ void qa_hash_insert(QaShared &S, int pid, int val) {
    unsigned slot = qa_hash(pid);
    for (;;) {
        const int old = atomicCAS(&S.hash_key[slot], 0, pid);
        if (old == 0 || old == pid) {
            S.hash_val[slot] = val;
            return;
        }
        slot = (slot + 1) & (QA_HASH - 1);
    }
}
// val of pid, or qa_pack(-1, 0, 0, 0) = 0 when the pid is unknown
__device__ __forceinline__ int qa_hash_find(const QaShared &S, int pid) {
    unsigned slot = qa_hash(pid);
    for (;;) {
        const int key = S.hash_key[slot];
        if (key == pid) return S.hash_val[slot];
        if (key == 0) return 0;
        slot = (slot + 1) & (QA_HASH - 1);
    }
}

__global__ void __launch_bounds__(QA_THREADS, QA_CTAS_PER_SM)
qa_kernel(const QaArgs A) {
    extern __shared__ __align__(16) unsigned char qa_smem[];
    QaShared &S = *reinterpret_cast<QaShared *>(qa_smem);
    double *qa = A.qa;
    double *my_pt_sum = A.scratch + static_cast<int64_t>(blockIdx.x)*QA_SCRATCH;    // [NSPEC][NPT]
    double *my_v2_num = my_pt_sum + ISS_QA_NSPEC*ISS_QA_NPT;                        // [NSPEC][NV2]
    for (int i = threadIdx.x; i < static_cast<int>(sizeof(QaShared)/4); i += blockDim.x)
        reinterpret_cast<unsigned int *>(qa_smem)[i] = 0u;
    __syncthreads();
    for (int i = threadIdx.x; i <= ISS_QA_NY; i += blockDim.x)
        S.y_edge[i] = sinh(-5.0 + i*(10.0/ISS_QA_NY));
    for (int i = threadIdx.x; i <= ISS_QA_NPHI; i += blockDim.x)
        sincos(-M_PI + i*(2.*M_PI/ISS_QA_NPHI), &S.phi_sin[i], &S.phi_cos[i]);
    // tracked pids first (charges 0), then the particle table (overwrites with the charges)
    if (threadIdx.x < A.npid && A.pids[threadIdx.x] != 0)
        qa_hash_insert(S, A.pids[threadIdx.x], qa_pack(threadIdx.x, 0, 0, 0));
    __syncthreads();
    if (A.dsp) {
        for (int j = threadIdx.x; j < A.ndsp; j += blockDim.x) {
            const iss_decay_species &d = A.dsp[j];
            int k = -1;
            for (int q = 0; q < A.npid; q++)
                if (A.pids[q] == d.pid) k = q;
            if (d.pid != 0) qa_hash_insert(S, d.pid, qa_pack(k, d.baryon, d.strange, d.charge));
        }
    }
    __syncthreads();

    // plain sums over all hadrons this thread sees (reduced once, at the end)
    double T[16];
#pragma unroll
    for (int i = 0; i < 16; i++) T[i] = 0.;
    double net[3] = {0., 0., 0.};
    int last_pid = 0, last_k = -1;
    double last_q[3] = {0., 0., 0.};

    for (int64_t ev = blockIdx.x; ev < A.nev; ev += gridDim.x) {
        const int64_t b = A.event_off[ev], e = A.event_off[ev + 1];
        double P[4] = {0., 0., 0., 0.};
        // 40-byte records as five 8-byte loads, the next record in flight while this one is binned
        const float2 *rec = reinterpret_cast<const float2 *>(A.hadrons);
        float2 nx[5];
        int64_t i = b + threadIdx.x;
        if (i < e) {
#pragma unroll
            for (int q = 0; q < 5; q++) nx[q] = __ldg(rec + 5*i + q);
        }
        for (; i < e; i += blockDim.x) {
            iss_hadron hd;
            hd.pid = __float_as_int(nx[0].x); hd.mass = nx[0].y;
            hd.E = nx[1].x; hd.px = nx[1].y;
            hd.py = nx[2].x; hd.pz = nx[2].y;
            hd.t = nx[3].x; hd.x = nx[3].y;
            hd.y = nx[4].x; hd.z = nx[4].y;
            if (i + blockDim.x < e) {
#pragma unroll
                for (int q = 0; q < 5; q++) nx[q] = __ldg(rec + 5*(i + blockDim.x) + q);
            }
            const double p[4] = {hd.E, hd.px, hd.py, hd.pz};
            const double inv_e = 1.0/p[0];
#pragma unroll
            for (int a = 0; a < 4; a++) {
                P[a] += p[a];
#pragma unroll
                for (int c = a; c < 4; c++) T[4*a + c] += p[a]*p[c]*inv_e;
            }
            if (hd.pid != last_pid) {
                last_pid = hd.pid;
                const int v = qa_hash_find(S, hd.pid);
                last_k = (v & 0xff) - 1;
                last_q[0] = static_cast<double>(static_cast<signed char>((v >> 8) & 0xff));
                last_q[1] = static_cast<double>(static_cast<signed char>((v >> 16) & 0xff));
                last_q[2] = static_cast<double>(static_cast<signed char>((v >> 24) & 0xff));
            }
            net[0] += last_q[0];
            net[1] += last_q[1];
            net[2] += last_q[2];
            const int k = last_k;
            if (k >= 0) {
                const double pT = sqrt(static_cast<double>(hd.px)*hd.px
                                       + static_cast<double>(hd.py)*hd.py);
                // Histogram::fill with bin_width = (5-0)/(100-1) (Histogram.cpp:12, 27-36)
                const double bw = 5.0/(ISS_QA_NPT - 1);
                const int ib = static_cast<int>(pT/bw);
                if (ib >= 0 && ib < ISS_QA_NPT) {
                    atomicAdd(&S.pt_evt[k][ib], 1);
                    atomicAdd(&my_pt_sum[k*ISS_QA_NPT + ib], pT);
                }
                const double mT2 = static_cast<double>(hd.mass)*hd.mass + pT*pT;
                // iy = floor((asinh(pz/mT) + 5)/0.1): float estimate, made exact against the
                // double-precision edges sinh(y_i)
                const double shy = hd.pz/sqrt(mT2);
                int iy = static_cast<int>(floorf((asinhf(static_cast<float>(shy)) + 5.f)
                                                 *(ISS_QA_NY/10.f)));
                iy = min(ISS_QA_NY, max(-1, iy));
                while (iy >= 0 && shy < S.y_edge[iy]) iy--;
                while (iy < ISS_QA_NY && shy >= S.y_edge[iy + 1]) iy++;
                if (iy >= 0 && iy < ISS_QA_NY) atomicAdd(&S.y_cnt[k][iy], 1u);
                // iphi = floor((atan2(py, px) + pi)/(2 pi/64)) clamped to [0, 63]: float estimate,
                // made exact with the sign of (cos, sin)(boundary) x (px, py)
                const double dpx = hd.px, dpy = hd.py;
                int iphi = ISS_QA_NPHI/2;           // atan2(0, 0) = 0
                if (dpx != 0. || dpy != 0.) {
                    iphi = static_cast<int>(floorf((atan2f(hd.py, hd.px) + 3.14159265f)
                                                   *(ISS_QA_NPHI/6.28318531f)));
                    iphi = min(ISS_QA_NPHI - 1, max(0, iphi));
                    while (iphi > 0 && dpy*S.phi_cos[iphi] - dpx*S.phi_sin[iphi] < 0.) iphi--;
                    while (iphi < ISS_QA_NPHI - 1
                           && dpy*S.phi_cos[iphi + 1] - dpx*S.phi_sin[iphi + 1] >= 0.) iphi++;
                }
                atomicAdd(&S.phi_cnt[k][iphi], 1u);
                const int iv = static_cast<int>(pT/(3.0/ISS_QA_NV2));
                if (iv < ISS_QA_NV2) {
                    const double c2 = (pT > 0.) ? (static_cast<double>(hd.px)*hd.px
                                                   - static_cast<double>(hd.py)*hd.py)/(pT*pT)
                                                : 0.;
                    atomicAdd(&my_v2_num[k*ISS_QA_NV2 + iv], c2);
                    atomicAdd(&S.v2_den[k][iv], 1u);
                }
                atomicAdd(&S.n_evt[k], 1);
            }
        }
        // per-event quantities: total four-momentum and the per-event bin contents
        for (int a = 0; a < 4; a++) {
            const double t = block_sum(P[a], S.red);
            if (threadIdx.x == 0) {
                S.P1[a] += t;
                S.P2[a] += t*t;
            }
        }
        __syncthreads();
        for (int i = threadIdx.x; i < A.npid*ISS_QA_NPT; i += blockDim.x) {
            const int k = i/ISS_QA_NPT, ib = i - k*ISS_QA_NPT;
            const unsigned long long c = S.pt_evt[k][ib];
            if (c) {
                S.pt_cnt[k][ib] += c;
                S.pt_sq[k][ib] += c*c;
                S.pt_evt[k][ib] = 0;
            }
        }
        if (threadIdx.x < A.npid) {
            const unsigned long long c = S.n_evt[threadIdx.x];
            S.n_tot[threadIdx.x] += c;
            S.n_sq[threadIdx.x] += c*c;
            S.n_evt[threadIdx.x] = 0;
        }
        if (threadIdx.x == 0) {
            S.n_had += e - b;
            S.n_evt_done++;
        }
        __syncthreads();
    }

    // flush: one global atomic per non-empty entry and CTA
    for (int a = 0; a < 4; a++)
        for (int c = a; c < 4; c++) {
            const double t = block_sum(T[4*a + c], S.red);
            if (threadIdx.x == 0 && t != 0.) {
                atomicAdd(&qa[9 + 4*a + c], t);
                if (c != a) atomicAdd(&qa[9 + 4*c + a], t);
            }
        }
    for (int a = 0; a < 3; a++) {
        const double t = block_sum(net[a], S.red);
        if (threadIdx.x == 0 && t != 0.) atomicAdd(&qa[26 + a], t);
    }
    if (threadIdx.x == 0 && S.n_evt_done > 0) {
        for (int a = 0; a < 4; a++) {
            atomicAdd(&qa[1 + a], S.P1[a]);
            atomicAdd(&qa[5 + a], S.P2[a]);
        }
        atomicAdd(&qa[0], static_cast<double>(S.n_evt_done));
        atomicAdd(&qa[25], static_cast<double>(S.n_had));
    }
    // (the CTA's own reductions at the L2 are complete before it reads them back, past the L1)
    __threadfence();
    __syncthreads();
    for (int k = 0; k < A.npid; k++) {
        double *blk = qa + ISS_QA_HEAD + static_cast<int64_t>(k)*ISS_QA_PER;
        for (int i = threadIdx.x; i < ISS_QA_NPT; i += blockDim.x) {
            if (S.pt_cnt[k][i]) {
                atomicAdd(&blk[i], static_cast<double>(S.pt_cnt[k][i]));
                atomicAdd(&blk[ISS_QA_NPT + i], __ldcg(&my_pt_sum[k*ISS_QA_NPT + i]));
                atomicAdd(&blk[2*ISS_QA_NPT + i], static_cast<double>(S.pt_sq[k][i]));
            }
        }
        for (int i = threadIdx.x; i < ISS_QA_NY; i += blockDim.x)
            if (S.y_cnt[k][i]) atomicAdd(&blk[3*ISS_QA_NPT + i], static_cast<double>(S.y_cnt[k][i]));
        for (int i = threadIdx.x; i < ISS_QA_NPHI; i += blockDim.x)
            if (S.phi_cnt[k][i])
                atomicAdd(&blk[3*ISS_QA_NPT + ISS_QA_NY + i], static_cast<double>(S.phi_cnt[k][i]));
        for (int i = threadIdx.x; i < ISS_QA_NV2; i += blockDim.x)
            if (S.v2_den[k][i]) {
                atomicAdd(&blk[3*ISS_QA_NPT + ISS_QA_NY + ISS_QA_NPHI + i],
                          __ldcg(&my_v2_num[k*ISS_QA_NV2 + i]));
                atomicAdd(&blk[3*ISS_QA_NPT + ISS_QA_NY + ISS_QA_NPHI + ISS_QA_NV2 + i],
                          static_cast<double>(S.v2_den[k][i]));
            }
        if (threadIdx.x == 0 && S.n_tot[k]) {
            atomicAdd(&blk[ISS_QA_PER - 2], static_cast<double>(S.n_tot[k]));
            atomicAdd(&blk[ISS_QA_PER - 1], static_cast<double>(S.n_sq[k]));
        }
    }
}

int run_qa(iss_handle *h, const int32_t *pids, int npid, int accumulate) {
    const int64_t nq = iss_cuda_qa_size();
    if (!h->d_qa) {
        ISS_CUDA_TRY(h, cudaMalloc(&h->d_qa, sizeof(double)*nq));
        accumulate = 0;
    }
    if (!accumulate) ISS_CUDA_TRY(h, cudaMemsetAsync(h->d_qa, 0, sizeof(double)*nq, h->stream));
    if (h->ndsp + npid > QA_HASH*3/4)
        ISS_FAIL(h, ISS_ERR_ARG, "particle table too large for the QA kernel's pid hash");
    QaArgs A;
    A.hadrons = h->decayed ? h->d_hadrons2 : h->d_hadrons;
    A.event_off = h->d_event_off;
    A.nev = h->ev_end - h->ev_begin;
    for (int i = 0; i < ISS_QA_NSPEC; i++) A.pids[i] = (i < npid) ? pids[i] : 0;
    A.npid = npid;
    A.dsp = h->d_dsp;
    A.ndsp = h->ndsp;
    A.qa = h->d_qa;
    int nsm = 148;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, h->device);
    const size_t smem = sizeof(QaShared);
    // per device and cheap: set on every call (a process may drive several devices)
    ISS_CUDA_TRY(h, cudaFuncSetAttribute(qa_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem)));
    int64_t grid = std::min<int64_t>(A.nev, static_cast<int64_t>(nsm)*QA_CTAS_PER_SM);
    if (grid < 1) grid = 1;
    ISS_ENSURE(h, h->d_qa_scratch, h->qa_scratch_bytes, sizeof(double)*QA_SCRATCH*static_cast<size_t>(nsm)*QA_CTAS_PER_SM);
    ISS_CUDA_TRY(h, cudaMemsetAsync(h->d_qa_scratch, 0, sizeof(double)*QA_SCRATCH*static_cast<size_t>(grid),
                                    h->stream));
    A.scratch = h->d_qa_scratch;
    {
        ScopedTimer t(h, ISS_T_QA);
        qa_kernel<<<static_cast<unsigned>(grid), QA_THREADS, smem, h->stream>>>(A); ISS_LAUNCHED(h);
    }
    ISS_CUDA_TRY(h, cudaGetLastError());
    return ISS_OK;      // asynchronous: iss_cuda_qa_fetch synchronises
}

}  // namespace iss
