// scan.cu -- exclusive prefix sum of int64 counts (multiplicity -> offsets).
// Integer prefix sums are associative, so any evaluation order gives the same
// bits; this replaces the implicit `push_back` bookkeeping of Hadron_list
// (FSSW.cpp:904-909, 1969-1996) by explicit offsets.
#include "iss_internal.cuh"

namespace iss {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_BLOCK = SCAN_THREADS*SCAN_ITEMS;   // 2048 elements per CTA

__device__ __forceinline__ int64_t block_exclusive_scan(int64_t v, int64_t *smem /*[8]*/,
                                                        int64_t &block_total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int64_t incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int64_t t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
    }
    if (lane == 31) smem[warp] = incl;
    __syncthreads();
    int64_t wbase = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < SCAN_THREADS/32; w++) {
        const int64_t x = smem[w];
        if (w < warp) wbase += x;
        tot += x;
    }
    block_total = tot;
    __syncthreads();
    return wbase + incl - v;
}

// phase 1: per-CTA local exclusive scan, CTA totals to `sums`
__global__ void __launch_bounds__(SCAN_THREADS)
scan_local_kernel(const int64_t *in, int64_t *out, int64_t *__restrict__ sums, int64_t n) {
    __shared__ int64_t smem[8];
    const int64_t base = static_cast<int64_t>(blockIdx.x)*SCAN_BLOCK + threadIdx.x*SCAN_ITEMS;
    int64_t v[SCAN_ITEMS];
    int64_t local = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) {
        v[i] = (base + i < n) ? in[base + i] : 0;
        local += v[i];
    }
    int64_t total;
    int64_t excl = block_exclusive_scan(local, smem, total);
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) {
        if (base + i < n) out[base + i] = excl;
        excl += v[i];
    }
    if (threadIdx.x == 0) sums[blockIdx.x] = total;
}

// phase 3: add scanned CTA offsets
__global__ void __launch_bounds__(SCAN_THREADS)
scan_add_kernel(int64_t *__restrict__ out, const int64_t *__restrict__ sums_scanned, int64_t n) {
    const int64_t base = static_cast<int64_t>(blockIdx.x)*SCAN_BLOCK + threadIdx.x*SCAN_ITEMS;
    const int64_t add = sums_scanned[blockIdx.x];
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++)
        if (base + i < n) out[base + i] += add;
}

__global__ void mail_post_kernel(const unsigned long long *__restrict__ src,
                                 unsigned long long *__restrict__ mail, int n) {
    if (threadIdx.x < n) mail[threadIdx.x] = src[threadIdx.x];
    __threadfence_system();
}

int mail_post(iss_handle *h, const void *d_src, int n, int slot) {
    if (!h->h_mail) {
        void *p = nullptr;
        ISS_CUDA_TRY(h, cudaHostAlloc(&p, sizeof(unsigned long long)*MAIL_WORDS, cudaHostAllocMapped));
        h->h_mail = static_cast<unsigned long long *>(p);
        void *d = nullptr;
        ISS_CUDA_TRY(h, cudaHostGetDevicePointer(&d, p, 0));
        h->d_mail = static_cast<unsigned long long *>(d);
    }
    mail_post_kernel<<<1, 32, 0, h->stream>>>(static_cast<const unsigned long long *>(d_src),
                                               h->d_mail + slot, n); ISS_LAUNCHED(h);
    ISS_CUDA_TRY(h, cudaGetLastError());
    return ISS_OK;
}

int ensure_mapped_event_offsets(iss_handle *h, int64_t n) {
    if (h->h_evoff && n <= h->evoff_mapped_cap) return ISS_OK;
    if (h->h_evoff) cudaFreeHost(h->h_evoff);
    h->h_evoff = nullptr;
    h->evoff_mapped_cap = n + n/2 + 1024;
    void *p = nullptr;
    ISS_CUDA_TRY(h, cudaHostAlloc(&p, sizeof(int64_t)*h->evoff_mapped_cap, cudaHostAllocMapped));
    h->h_evoff = static_cast<int64_t *>(p);
    void *d = nullptr;
    ISS_CUDA_TRY(h, cudaHostGetDevicePointer(&d, p, 0));
    h->d_evoff_mapped = static_cast<int64_t *>(d);
    return ISS_OK;
}

static int scan_rec(iss_handle *h, const int64_t *d_in, int64_t *d_out, int64_t n,
                    int64_t *d_tmp, int64_t tmp_elems) {
    const int64_t nblk = (n + SCAN_BLOCK - 1)/SCAN_BLOCK;
    if (nblk > tmp_elems) ISS_FAIL(h, ISS_ERR_NOMEM, "scan scratch too small");
    int64_t *sums = d_tmp;
    scan_local_kernel<<<static_cast<unsigned>(nblk), SCAN_THREADS, 0, h->stream>>>(d_in, d_out,
                                                                                   sums, n); ISS_LAUNCHED(h);
    if (nblk > 1) {
        int rc = scan_rec(h, sums, sums, nblk, d_tmp + nblk, tmp_elems - nblk);
        if (rc) return rc;
        scan_add_kernel<<<static_cast<unsigned>(nblk), SCAN_THREADS, 0, h->stream>>>(d_out, sums,
                                                                                     n); ISS_LAUNCHED(h);
    }
    ISS_CUDA_TRY(h, cudaGetLastError());
    return ISS_OK;
}

// d_out[i] = sum_{j<i} d_in[j] for i in [0, n]; d_out has n+1 entries (last = total).
// d_in must have n+1 readable entries with d_in[n] ignored (treated as 0 by the caller
// zero-filling it).  If h_total != nullptr the total is copied back (stream synchronised).
int device_exclusive_scan_i64(iss_handle *h, const int64_t *d_in, int64_t *d_out, int64_t n,
                              int64_t *h_total) {
    const int64_t n1 = n + 1;
    int64_t need = 0;
    for (int64_t m = n1; m > 1;) {
        m = (m + SCAN_BLOCK - 1)/SCAN_BLOCK;
        need += m;
        if (m == 1) break;
    }
    need += 2;
    if (h->scan_tmp_bytes < sizeof(int64_t)*need) {
        if (h->d_scan_tmp) cudaFree(h->d_scan_tmp);
        h->d_scan_tmp = nullptr;
        h->scan_tmp_bytes = sizeof(int64_t)*(need + 64);
        ISS_CUDA_TRY(h, cudaMalloc(&h->d_scan_tmp, h->scan_tmp_bytes));
    }
    int rc = scan_rec(h, d_in, d_out, n1, static_cast<int64_t *>(h->d_scan_tmp),
                      static_cast<int64_t>(h->scan_tmp_bytes/sizeof(int64_t)));
    if (rc) return rc;
    if (h_total) {
        rc = mail_post(h, d_out + n, 1, 8);
        if (rc) return rc;
        ISS_CUDA_TRY(h, cudaStreamSynchronize(h->stream));
        *h_total = static_cast<int64_t>(*reinterpret_cast<volatile unsigned long long *>(h->h_mail + 8));
    }
    return ISS_OK;
}

}  // namespace iss
