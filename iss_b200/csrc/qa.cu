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
    const int32_t *sorted_pid;
    const int32_t *sorted_idx;
    int ndsp;
    double *qa;
};

constexpr int QA_THREADS = 256;

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

__global__ void __launch_bounds__(QA_THREADS)
qa_kernel(const QaArgs A) {
    __shared__ int pt_cnt[ISS_QA_NSPEC][ISS_QA_NPT];
    __shared__ int n_cnt[ISS_QA_NSPEC];
    __shared__ double red[QA_THREADS/32];
    double *qa = A.qa;
    for (int64_t ev = blockIdx.x; ev < A.nev; ev += gridDim.x) {
        for (int i = threadIdx.x; i < ISS_QA_NSPEC*ISS_QA_NPT; i += blockDim.x)
            (&pt_cnt[0][0])[i] = 0;
        if (threadIdx.x < ISS_QA_NSPEC) n_cnt[threadIdx.x] = 0;
        __syncthreads();
        const int64_t b = A.event_off[ev], e = A.event_off[ev + 1];
        double P[4] = {0., 0., 0., 0.};
        double T[16];
#pragma unroll
        for (int i = 0; i < 16; i++) T[i] = 0.;
        double net[3] = {0., 0., 0.};
        for (int64_t i = b + threadIdx.x; i < e; i += blockDim.x) {
            const iss_hadron hd = A.hadrons[i];
            const double p[4] = {hd.E, hd.px, hd.py, hd.pz};
#pragma unroll
            for (int a = 0; a < 4; a++) {
                P[a] += p[a];
#pragma unroll
                for (int c = 0; c < 4; c++) T[4*a + c] += p[a]*p[c]/p[0];
            }
            if (A.dsp) {
                int lo = 0, hi = A.ndsp - 1;
                while (lo <= hi) {
                    const int mid = (lo + hi) >> 1;
                    const int v = __ldg(&A.sorted_pid[mid]);
                    if (v == hd.pid) {
                        const iss_decay_species &s = A.dsp[__ldg(&A.sorted_idx[mid])];
                        net[0] += s.baryon;
                        net[1] += s.strange;
                        net[2] += s.charge;
                        break;
                    }
                    if (v < hd.pid) lo = mid + 1; else hi = mid - 1;
                }
            }
            int k = -1;
            for (int j = 0; j < A.npid; j++)
                if (A.pids[j] == hd.pid) k = j;
            if (k >= 0) {
                double *blk = qa + ISS_QA_HEAD + static_cast<int64_t>(k)*ISS_QA_PER;
                const double pT = sqrt(static_cast<double>(hd.px)*hd.px
                                       + static_cast<double>(hd.py)*hd.py);
                // Histogram::fill with bin_width = (5-0)/(100-1) (Histogram.cpp:12, 27-36)
                const double bw = 5.0/(ISS_QA_NPT - 1);
                const int ib = static_cast<int>(pT/bw);
                if (ib >= 0 && ib < ISS_QA_NPT) {
                    atomicAdd(&pt_cnt[k][ib], 1);
                    atomicAdd(&blk[ISS_QA_NPT + ib], pT);
                }
                const double mT2 = static_cast<double>(hd.mass)*hd.mass + pT*pT;
                const double y = asinh(hd.pz/sqrt(mT2));
                const int iy = static_cast<int>(floor((y + 5.0)/(10.0/ISS_QA_NY)));
                if (iy >= 0 && iy < ISS_QA_NY) atomicAdd(&blk[3*ISS_QA_NPT + iy], 1.0);
                const double phi = atan2(static_cast<double>(hd.py), static_cast<double>(hd.px));
                int iphi = static_cast<int>(floor((phi + M_PI)/(2.*M_PI/ISS_QA_NPHI)));
                iphi = min(ISS_QA_NPHI - 1, max(0, iphi));
                atomicAdd(&blk[3*ISS_QA_NPT + ISS_QA_NY + iphi], 1.0);
                const int iv = static_cast<int>(pT/(3.0/ISS_QA_NV2));
                if (iv < ISS_QA_NV2) {
                    const double c2 = (pT > 0.) ? (static_cast<double>(hd.px)*hd.px
                                                   - static_cast<double>(hd.py)*hd.py)/(pT*pT)
                                                : 0.;
                    atomicAdd(&blk[3*ISS_QA_NPT + ISS_QA_NY + ISS_QA_NPHI + iv], c2);
                    atomicAdd(&blk[3*ISS_QA_NPT + ISS_QA_NY + ISS_QA_NPHI + ISS_QA_NV2 + iv], 1.0);
                }
                atomicAdd(&n_cnt[k], 1);
            }
        }
        __syncthreads();
        // per-event quantities
        for (int a = 0; a < 4; a++) {
            const double t = block_sum(P[a], red);
            if (threadIdx.x == 0) {
                atomicAdd(&qa[1 + a], t);
                atomicAdd(&qa[5 + a], t*t);
            }
        }
        for (int a = 0; a < 16; a++) {
            const double t = block_sum(T[a], red);
            if (threadIdx.x == 0) atomicAdd(&qa[9 + a], t);
        }
        for (int a = 0; a < 3; a++) {
            const double t = block_sum(net[a], red);
            if (threadIdx.x == 0) atomicAdd(&qa[26 + a], t);
        }
        if (threadIdx.x == 0) {
            atomicAdd(&qa[0], 1.0);
            atomicAdd(&qa[25], static_cast<double>(e - b));
        }
        for (int i = threadIdx.x; i < A.npid*ISS_QA_NPT; i += blockDim.x) {
            const int k = i/ISS_QA_NPT, ib = i - k*ISS_QA_NPT;
            const int c = pt_cnt[k][ib];
            if (c) {
                double *blk = qa + ISS_QA_HEAD + static_cast<int64_t>(k)*ISS_QA_PER;
                atomicAdd(&blk[ib], static_cast<double>(c));
                atomicAdd(&blk[2*ISS_QA_NPT + ib], static_cast<double>(c)*c);
            }
        }
        if (threadIdx.x < A.npid) {
            const int k = threadIdx.x;
            const double c = n_cnt[k];
            double *blk = qa + ISS_QA_HEAD + static_cast<int64_t>(k)*ISS_QA_PER;
            atomicAdd(&blk[ISS_QA_PER - 2], c);
            atomicAdd(&blk[ISS_QA_PER - 1], c*c);
        }
        __syncthreads();
    }
}

int run_qa(iss_handle *h, const int32_t *pids, int npid, int accumulate) {
    const int64_t nq = iss_cuda_qa_size();
    if (!h->d_qa) {
        ISS_CUDA_TRY(h, cudaMalloc(&h->d_qa, sizeof(double)*nq));
        accumulate = 0;
    }
    if (!accumulate) ISS_CUDA_TRY(h, cudaMemsetAsync(h->d_qa, 0, sizeof(double)*nq, h->stream));
    QaArgs A;
    A.hadrons = h->decayed ? h->d_hadrons2 : h->d_hadrons;
    A.event_off = h->d_event_off;
    A.nev = h->ev_end - h->ev_begin;
    for (int i = 0; i < ISS_QA_NSPEC; i++) A.pids[i] = (i < npid) ? pids[i] : 0;
    A.npid = npid;
    A.dsp = h->d_dsp;
    A.sorted_pid = h->d_sorted_pid;
    A.sorted_idx = h->d_sorted_idx;
    A.ndsp = h->ndsp;
    A.qa = h->d_qa;
    int nsm = 148;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, h->device);
    int64_t grid = std::min<int64_t>(A.nev, static_cast<int64_t>(nsm)*8);
    if (grid < 1) grid = 1;
    {
        ScopedTimer t(h, ISS_T_QA);
        qa_kernel<<<static_cast<unsigned>(grid), QA_THREADS, 0, h->stream>>>(A); ISS_LAUNCHED(h);
    }
    ISS_CUDA_TRY(h, cudaGetLastError());
    ISS_CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return ISS_OK;
}

}  // namespace iss
