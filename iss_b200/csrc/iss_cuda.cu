// iss_cuda.cu -- C-ABI entry points (include/iss_cuda.h): handle life-time, uploads,
// orchestration of the kernels in yields.cu / sampler.cu / decay.cu / qa.cu and the
// device->host accessors that replace FSSW's Hadron_list getters (FSSW.h:166-184).
#include <cmath>
#include <algorithm>
#include <cstring>
#include <new>
#include <utility>

#include "iss_internal.cuh"

using namespace iss;

namespace {

// SoA [ISS_NFIELD][ncell_pad] -> AoS [ncell][CELL_STRIDE] with t = tau cosh(eta), z = tau sinh(eta)
// precomputed in double from the float fields exactly as FSSW::add_one_sampled_particle does for
// eta_s = cell eta (FSSW.cpp:1981-1982).
__global__ void build_cells_kernel(const float *__restrict__ soa, int64_t ncell, int64_t ncell_pad,
                                   float *__restrict__ cells, float4 *__restrict__ thermo) {
    const int64_t c = static_cast<int64_t>(blockIdx.x)*blockDim.x + threadIdx.x;
    if (c >= ncell) return;
    float rec[CELL_STRIDE];
#pragma unroll
    for (int k = 0; k < ISS_NFIELD; k++) rec[k] = soa[static_cast<int64_t>(k)*ncell_pad + c];
    const double tau = rec[ISS_F_TAU], eta = rec[ISS_F_ETA];
    rec[CELL_T] = static_cast<float>(tau*cosh(eta));
    rec[CELL_Z] = static_cast<float>(tau*sinh(eta));
    rec[30] = 0.f;
    rec[31] = 0.f;
    thermo[c] = make_float4(rec[ISS_F_T], rec[ISS_F_MUB], rec[ISS_F_MUS], rec[ISS_F_MUQ]);
    float4 *dst = reinterpret_cast<float4 *>(cells + c*CELL_STRIDE);
#pragma unroll
    for (int k = 0; k < CELL_STRIDE/4; k++)
        dst[k] = make_float4(rec[4*k], rec[4*k + 1], rec[4*k + 2], rec[4*k + 3]);
}

__global__ void fp64_fma_kernel(double *out, int iters) {
    double a0 = threadIdx.x*1e-9, a1 = a0 + 1., a2 = a0 + 2., a3 = a0 + 3.;
    double a4 = a0 + 4., a5 = a0 + 5., a6 = a0 + 6., a7 = a0 + 7.;
    const double b = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; i++) {
        a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
        a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
    }
    out[static_cast<size_t>(blockIdx.x)*blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

void free_surface(iss_handle *h) {
    cudaFree(h->d_surf); h->d_surf = nullptr; h->surf_bytes = 0;
    cudaFree(h->d_cells); h->d_cells = nullptr; h->cells_bytes = 0;
    cudaFree(h->d_cellcoef); h->d_cellcoef = nullptr; h->coef_bytes = 0;
    cudaFree(h->d_stage); h->d_stage = nullptr; h->stage_bytes = 0;
    cudaFree(h->d_thermo); h->d_thermo = nullptr; h->thermo_bytes = 0;
    cudaFree(h->d_yields); h->d_yields = nullptr; h->yields_bytes = 0;
    cudaFree(h->d_cdf); h->d_cdf = nullptr; h->cdf_bytes = 0;
    cudaFree(h->d_tilesum); h->d_tilesum = nullptr; h->tilesum_bytes = 0;
    cudaFree(h->d_tilebase); h->d_tilebase = nullptr; h->tilebase_bytes = 0;
    cudaFree(h->d_total); h->d_total = nullptr; h->total_bytes = 0;
    cudaFree(h->d_cdflev); h->d_cdflev = nullptr; h->cdflev_bytes = 0;
    cudaFree(h->d_guide); h->d_guide = nullptr; h->guide_bytes = 0; h->guide_M = 0;
    cudaFree(h->d_tilesum_g); h->d_tilesum_g = nullptr; h->tilesum_g_bytes = 0;
    cudaFree(h->d_tilebase_g); h->d_tilebase_g = nullptr; h->tilebase_g_bytes = 0;
    cudaFree(h->d_cdflev_g); h->d_cdflev_g = nullptr; h->cdflev_g_bytes = 0;
    h->have_yields = false;
    h->have_local_yields = false;
    h->cellrec_valid = false;
    h->have_batch = false;
}

// [ncell][28] staging -> SoA [28][ncell_pad] and the sampler's AoS [ncell][32]; one thread per
// cell reads its 112-byte record with 16-byte loads
__global__ void unpack_cells_kernel(const float *__restrict__ stage, int64_t first, int64_t end,
                                    int64_t ncell_pad, float *__restrict__ soa,
                                    float *__restrict__ cells, float4 *__restrict__ thermo) {
    const int64_t c = first + static_cast<int64_t>(blockIdx.x)*blockDim.x + threadIdx.x;
    if (c >= end) return;
    float rec[CELL_STRIDE];
    const float4 *src = reinterpret_cast<const float4 *>(stage + c*ISS_NFIELD);
#pragma unroll
    for (int k = 0; k < ISS_NFIELD/4; k++) {
        const float4 v = __ldg(src + k);
        rec[4*k] = v.x; rec[4*k + 1] = v.y; rec[4*k + 2] = v.z; rec[4*k + 3] = v.w;
    }
#pragma unroll
    for (int k = 0; k < ISS_NFIELD; k++) soa[static_cast<int64_t>(k)*ncell_pad + c] = rec[k];
    const double tau = rec[ISS_F_TAU], eta = rec[ISS_F_ETA];
    rec[CELL_T] = static_cast<float>(tau*cosh(eta));
    rec[CELL_Z] = static_cast<float>(tau*sinh(eta));
    rec[30] = 0.f;
    rec[31] = 0.f;
    thermo[c] = make_float4(rec[ISS_F_T], rec[ISS_F_MUB], rec[ISS_F_MUS], rec[ISS_F_MUQ]);
    float4 *dst = reinterpret_cast<float4 *>(cells + c*CELL_STRIDE);
#pragma unroll
    for (int k = 0; k < CELL_STRIDE/4; k++)
        dst[k] = make_float4(rec[4*k], rec[4*k + 1], rec[4*k + 2], rec[4*k + 3]);
}

int prepare_surface_buffers(iss_handle *h, int64_t ncell) {
    if (ncell >= (int64_t(1) << 31)) ISS_FAIL(h, ISS_ERR_ARG, "ncell must be < 2^31");
    h->ncell = ncell;
    h->ncell_lrf = ncell;
    h->legacy = false;
    h->ntile = (ncell + TILE - 1)/TILE;
    h->ncell_pad = h->ntile*TILE;
    h->have_yields = false;
    h->have_local_yields = false;
    h->have_batch = false;
    h->chunk = false;           // a new surface is a whole surface until declared a chunk
    ISS_ENSURE(h, h->d_surf, h->surf_bytes, sizeof(float)*ISS_NFIELD*h->ncell_pad);
    ISS_ENSURE(h, h->d_cells, h->cells_bytes, sizeof(float)*CELL_STRIDE*ncell);
    ISS_ENSURE(h, h->d_thermo, h->thermo_bytes, sizeof(float4)*ncell);
    return ISS_OK;
}

int upload_doubles(iss_handle *h, double **dptr, const double *src, size_t n) {
    if (*dptr) cudaFree(*dptr);
    *dptr = nullptr;
    ISS_CUDA_TRY(h, cudaMalloc(dptr, sizeof(double)*n));
    ISS_CUDA_TRY(h, cudaMemcpyAsync(*dptr, src, sizeof(double)*n, cudaMemcpyHostToDevice,
                                    h->stream));
    ISS_CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return ISS_OK;
}

}  // namespace

namespace iss {

// Surface-chunk mode: the ranks' tile sums ([nspecies][rank_ntile[r]] each, in rank = cell order)
// into the columns of the [nspecies][g_ntile] table of the whole surface, then the fixed-order
// combination every rank evaluates (run_yields_finish).
int chunk_combine_tile_sums(iss_handle *h, const double *const *rank_tilesums, const int64_t *rank_ntile,
                            int32_t nranks, int on_device, double *dN_species_host) {
    int64_t sum = 0;
    bool mine = false;
    for (int r = 0; r < nranks; r++) {
        if (rank_ntile[r] < 0 || !rank_tilesums[r]) return ISS_ERR_ARG;
        if (sum == h->chunk_tile_begin && rank_ntile[r] == h->ntile) mine = true;
        sum += rank_ntile[r];
    }
    if (sum != h->g_ntile || !mine)
        ISS_FAIL(h, ISS_ERR_ARG, "the ranks' tile counts do not add up to the surface declared by "
                                 "iss_cuda_set_surface_chunk, or this handle's chunk is not among them");
    const int64_t ns = h->nspecies;
    ISS_ENSURE(h, h->d_tilesum_g, h->tilesum_g_bytes, sizeof(double)*ns*h->g_ntile);
    int64_t t0 = 0;
    for (int r = 0; r < nranks; r++) {
        // [ns][rank_ntile[r]] -> columns [t0, t0 + rank_ntile[r]) of [ns][g_ntile]
        if (rank_ntile[r] > 0)
            ISS_CUDA_TRY(h, cudaMemcpy2DAsync(h->d_tilesum_g + t0, sizeof(double)*h->g_ntile,
                                              rank_tilesums[r], sizeof(double)*rank_ntile[r],
                                              sizeof(double)*rank_ntile[r], ns,
                                              on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice,
                                              h->stream));
        t0 += rank_ntile[r];
    }
    int rc = run_yields_finish(h);
    if (rc) return rc;
    if (dN_species_host) memcpy(dN_species_host, h->h_total.data(), sizeof(double)*h->nspecies);
    return ISS_OK;
}

}  // namespace iss

extern "C" {

int iss_cuda_create(int device, iss_handle **out) {
    if (!out) return ISS_ERR_ARG;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev)
        return ISS_ERR_CUDA;
    if (cudaSetDevice(device) != cudaSuccess) return ISS_ERR_CUDA;
    iss_handle *h = new (std::nothrow) iss_handle();
    if (!h) return ISS_ERR_NOMEM;
    h->device = device;
    if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete h;
        return ISS_ERR_CUDA;
    }
    h->own_stream = true;
    // the kernels exist for sm_100a only: a device that cannot run them is refused here, not at
    // the first launch
    cudaFuncAttributes fa;
    bool ok = cudaFuncGetAttributes(&fa, fp64_fma_kernel) == cudaSuccess;
    ok = ok && cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&h->batch_ready, cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&h->copy_done[0], cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&h->copy_done[1], cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&h->copy_done2, cudaEventDisableTiming) == cudaSuccess;
    if (!ok) {
        cudaGetLastError();     // clear the sticky-free error state of this thread
        iss_cuda_destroy(h);
        return ISS_ERR_CUDA;
    }
    *out = h;
    return ISS_OK;
}

int iss_cuda_destroy(iss_handle *h) {
    if (!h) return ISS_OK;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    if (h->copy_stream) { cudaStreamSynchronize(h->copy_stream); cudaStreamDestroy(h->copy_stream); }
    if (h->batch_ready) cudaEventDestroy(h->batch_ready);
    for (int b = 0; b < 2; b++) if (h->copy_done[b]) cudaEventDestroy(h->copy_done[b]);
    if (h->copy_done2) cudaEventDestroy(h->copy_done2);
    free_surface(h);
    cudaFree(h->d_species);
    cudaFree(h->d_bessel); cudaFree(h->d_expint); cudaFree(h->d_sf4); cudaFree(h->d_combos); cudaFree(h->d_ce); cudaFree(h->d_mom22);
    cudaFree(h->d_mom14); cudaFree(h->d_kappa);
    cudaFree(h->d_lab); cudaFree(h->d_labrec); cudaFree(h->d_spec_part); cudaFree(h->d_spec_out);
    cudaFree(h->d_spec_tab); cudaFree(h->d_ingest);
    for (int r = 0; r < 6; r++) cudaFree(h->d_momtab[r]);
    cudaFree(h->d_dsp); cudaFree(h->d_dch); cudaFree(h->d_sorted_pid); cudaFree(h->d_sorted_idx);
    cudaFree(h->d_lambda); cudaFree(h->d_pmode);
    cudaFree(h->d_mult); cudaFree(h->d_off_out); cudaFree(h->d_off_work);
    cudaFree(h->d_hadbuf[0]); cudaFree(h->d_hadbuf[1]); cudaFree(h->d_hadrons2); cudaFree(h->d_event_off);
    cudaFree(h->d_tasks); cudaFree(h->d_sampler_args); cudaFree(h->d_hints);
    cudaFree(h->d_task_slot); cudaFree(h->d_tasks_unsorted);
    cudaFree(h->d_cellcnt); cudaFree(h->d_cellrec);
    cudaFree(h->d_counters); cudaFree(h->d_decay_cnt); cudaFree(h->d_scan_tmp);
    cudaFree(h->d_qa); cudaFree(h->d_trace);
    cudaFree(h->d_qa_scratch); cudaFree(h->d_own); cudaFree(h->d_wlist); cudaFree(h->d_ownmask); cudaFree(h->d_tilesum_all);
    cudaFree(h->d_legpos); cudaFree(h->d_legcoef); cudaFree(h->d_zx); cudaFree(h->d_zy);
    cudaFree(h->d_lambert); cudaFree(h->d_legmax); cudaFree(h->d_bulk0);
    if (h->h_mail) cudaFreeHost(h->h_mail);
    if (h->h_evoff) cudaFreeHost(h->h_evoff);
    for (auto &sp : h->spans) { cudaEventDestroy(sp.a); cudaEventDestroy(sp.b); }
    for (auto e : h->ev_pool) cudaEventDestroy(e);
    if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return ISS_OK;
}

const char *iss_cuda_last_error(const iss_handle *h) {
    return h ? h->err.c_str() : "null handle";
}

int iss_cuda_set_stream(iss_handle *h, void *cuda_stream) {
    if (!h) return ISS_ERR_ARG;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
    h->stream = static_cast<cudaStream_t>(cuda_stream);
    h->own_stream = false;
    return ISS_OK;
}

int iss_cuda_synchronize(iss_handle *h) {
    if (!h) return ISS_ERR_ARG;
    ISS_CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return ISS_OK;
}

int iss_cuda_upload_surface(iss_handle *h, const float *const soa[ISS_NFIELD], int64_t ncell) {
    if (!h || !soa || ncell <= 0) return ISS_ERR_ARG;
    cudaSetDevice(h->device);
    int rc = prepare_surface_buffers(h, ncell);
    if (rc) return rc;
    ISS_CUDA_TRY(h, cudaMemsetAsync(h->d_surf, 0, sizeof(float)*ISS_NFIELD*h->ncell_pad,
                                    h->stream));
    for (int k = 0; k < ISS_NFIELD; k++) {
        if (!soa[k]) ISS_FAIL(h, ISS_ERR_ARG, "null field pointer");
        ISS_CUDA_TRY(h, cudaMemcpyAsync(h->d_surf + static_cast<int64_t>(k)*h->ncell_pad, soa[k],
                                        sizeof(float)*ncell, cudaMemcpyHostToDevice, h->stream));
    }
    build_cells_kernel<<<static_cast<unsigned>((ncell + 127)/128), 128, 0, h->stream>>>(
        h->d_surf, ncell, h->ncell_pad, h->d_cells, h->d_thermo); ISS_LAUNCHED(h);
    ISS_CUDA_TRY(h, cudaGetLastError());
    ISS_CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return ISS_OK;
}

int iss_cuda_upload_surface_aos(iss_handle *h, const float *cells, int64_t ncell) {
    if (!h || !cells || ncell <= 0) return ISS_ERR_ARG;
    cudaSetDevice(h->device);
    int rc = prepare_surface_buffers(h, ncell);
    if (rc) return rc;
    ISS_ENSURE(h, h->d_stage, h->stage_bytes, sizeof(float)*ISS_NFIELD*ncell);
    ISS_CUDA_TRY(h, cudaMemcpyAsync(h->d_stage, cells, sizeof(float)*ISS_NFIELD*ncell,
                                    cudaMemcpyHostToDevice, h->stream));
    if (h->ncell_pad > ncell)   // padding cells read as zeros
        for (int k = 0; k < ISS_NFIELD; k++)
            ISS_CUDA_TRY(h, cudaMemsetAsync(h->d_surf + static_cast<int64_t>(k)*h->ncell_pad + ncell,
                                            0, sizeof(float)*(h->ncell_pad - ncell), h->stream));
    unpack_cells_kernel<<<static_cast<unsigned>((ncell + 127)/128), 128, 0, h->stream>>>(
        h->d_stage, 0, ncell, h->ncell_pad, h->d_surf, h->d_cells, h->d_thermo); ISS_LAUNCHED(h);
    ISS_CUDA_TRY(h, cudaGetLastError());
    // the caller's buffer may be reused as soon as this returns
    ISS_CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return ISS_OK;
}

int iss_cuda_upload_surface_aos_part(iss_handle *h, const float *cells_part, int64_t first,
                                     int64_t n, int64_t ncell_total) {
    if (!h || !cells_part || first < 0 || n <= 0 || first + n > ncell_total) return ISS_ERR_ARG;
    cudaSetDevice(h->device);
    if (first == 0) {
        int rc = prepare_surface_buffers(h, ncell_total);
        if (rc) return rc;
        ISS_ENSURE(h, h->d_stage, h->stage_bytes, sizeof(float)*ISS_NFIELD*ncell_total);
        if (h->ncell_pad > ncell_total)   // padding cells read as zeros
            for (int k = 0; k < ISS_NFIELD; k++)
                ISS_CUDA_TRY(h, cudaMemsetAsync(h->d_surf + static_cast<int64_t>(k)*h->ncell_pad + ncell_total,
                                                0, sizeof(float)*(h->ncell_pad - ncell_total), h->stream));
    } else if (h->ncell != ncell_total || !h->d_stage) {
        ISS_FAIL(h, ISS_ERR_STATE, "surface parts must start with first = 0 and keep ncell_total");
    }
    // copy and transposition of this part queue behind each other on the stream; the host packs
    // the next part meanwhile
    ISS_CUDA_TRY(h, cudaMemcpyAsync(h->d_stage + first*ISS_NFIELD, cells_part, sizeof(float)*ISS_NFIELD*n,
                                    cudaMemcpyHostToDevice, h->stream));
    unpack_cells_kernel<<<static_cast<unsigned>((n + 127)/128), 128, 0, h->stream>>>(
        h->d_stage, first, first + n, h->ncell_pad, h->d_surf, h->d_cells, h->d_thermo); ISS_LAUNCHED(h);
    ISS_CUDA_TRY(h, cudaGetLastError());
    if (first + n == ncell_total) ISS_CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return ISS_OK;
}

int iss_cuda_upload_species(iss_handle *h, const iss_species *species, int32_t nspecies) {
    if (!h || !species || nspecies <= 0 || nspecies > MAX_SPECIES) return ISS_ERR_ARG;
    cudaSetDevice(h->device);
    h->h_species.assign(species, species + nspecies);
    std::vector<DeviceSpecies> ds(nspecies);
    std::vector<int4> combos;
    for (int i = 0; i < nspecies; i++) {
        const iss_species &p = species[i];
        DeviceSpecies &d = ds[i];
        d.mass = p.mass;
        d.mass2 = p.mass*p.mass;
        d.inv_mass = 1.0/p.mass;
        d.pid = p.pid;
        d.gspin = static_cast<int16_t>(p.gspin);
        d.baryon = static_cast<int16_t>(p.baryon);
        d.strange = static_cast<int16_t>(p.strange);
        d.charge = static_cast<int16_t>(p.charge);
        d.sign = static_cast<int16_t>(p.sign);
        d.trunc10_mass = (p.mass < 0.7) ? 1 : 0;    // FSSW.cpp:741
        d.decay_idx = p.decay_idx;
        int c = 0;
        for (; c < static_cast<int>(combos.size()); c++)
            if (combos[c].x == p.baryon && combos[c].y == p.strange && combos[c].z == p.charge) break;
        if (c == static_cast<int>(combos.size())) combos.push_back(make_int4(p.baryon, p.strange, p.charge, 0));
        d.combo = c;
    }
    if (combos.size() > 200) ISS_FAIL(h, ISS_ERR_ARG, "more than 200 distinct (B,S,Q) combinations");
    if (h->d_combos) cudaFree(h->d_combos);
    h->d_combos = nullptr;
    ISS_CUDA_TRY(h, cudaMalloc(&h->d_combos, sizeof(int4)*combos.size()));
    ISS_CUDA_TRY(h, cudaMemcpyAsync(h->d_combos, combos.data(), sizeof(int4)*combos.size(),
                                    cudaMemcpyHostToDevice, h->stream));
    h->ncombo = static_cast<int>(combos.size());
    if (h->d_species) cudaFree(h->d_species);
    h->d_species = nullptr;
    ISS_CUDA_TRY(h, cudaMalloc(&h->d_species, sizeof(DeviceSpecies)*nspecies));
    ISS_CUDA_TRY(h, cudaMemcpyAsync(h->d_species, ds.data(), sizeof(DeviceSpecies)*nspecies,
                                    cudaMemcpyHostToDevice, h->stream));
    ISS_CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    h->nspecies = nspecies;
    h->have_yields = false;
    return ISS_OK;
}

int iss_cuda_upload_table(iss_handle *h, int32_t kind, const double *data, int64_t n0, int64_t n1,
                          const double *grid4) {
    if (!h || !data || n0 <= 0) return ISS_ERR_ARG;
    cudaSetDevice(h->device);
    h->have_yields = false;
    switch (kind) {
    case ISS_TABLE_BESSEL_K:
    case ISS_TABLE_EXPINT: {
        if (!grid4) return ISS_ERR_ARG;
        SfGrid g;
        g.x_min = grid4[0];
        g.dx = grid4[1];
        g.n = static_cast<int>(n0);
        g.x_max_minus_dx = grid4[2];    // = sf_x_max - sf_dx of the caller
        h->sf = g;
        if (h->d_sf4) { cudaFree(h->d_sf4); h->d_sf4 = nullptr; }
        if (kind == ISS_TABLE_BESSEL_K) return upload_doubles(h, &h->d_bessel, data, n0*3);
        return upload_doubles(h, &h->d_expint, data, n0*9);
    }
    case ISS_TABLE_CE:
        if (n0 != n1) ISS_FAIL(h, ISS_ERR_ARG, "CE table must be square (reference indexes with one length)");
        h->ce_ne = static_cast<int>(n0);
        h->ce_nb = static_cast<int>(n1);
        return upload_doubles(h, &h->d_ce, data, n0*n1*5);
    case ISS_TABLE_MOM22:
        if (n0 != n1) ISS_FAIL(h, ISS_ERR_ARG, "22-moment table must be square");
        h->ce_ne = static_cast<int>(n0);
        h->ce_nb = static_cast<int>(n1);
        return upload_doubles(h, &h->d_mom22, data, n0*n1*8);
    case ISS_TABLE_MOM14:
        if (!grid4 || n1 <= 0) return ISS_ERR_ARG;
        h->g14 = Grid2D{grid4[0], grid4[1], grid4[2], grid4[3], static_cast<int>(n0),
                        static_cast<int>(n1)};
        return upload_doubles(h, &h->d_mom14, data, 3*n0*n1);
    case ISS_TABLE_BULK14: {
        // rows "T[1/fm] B0 D0 E0" -> four columns (Table::interp works column-wise)
        if (n1 != 4 || n0 < 4) ISS_FAIL(h, ISS_ERR_ARG, "14-moment bulk table: n0 >= 4 rows of 4 numbers");
        std::vector<double> cols(static_cast<size_t>(4)*n0);
        for (int64_t i = 0; i < n0; i++)
            for (int c = 0; c < 4; c++) cols[c*n0 + i] = data[i*4 + c];
        h->nbulk0 = static_cast<int>(n0);
        return upload_doubles(h, &h->d_bulk0, cols.data(), cols.size());
    }
    case ISS_TABLE_KAPPA_B:
        if (!grid4 || n1 <= 0) return ISS_ERR_ARG;
        h->gk = Grid2D{grid4[0], grid4[1], grid4[2], grid4[3], static_cast<int>(n0),
                       static_cast<int>(n1)};
        return upload_doubles(h, &h->d_kappa, data, n0*n1);
    default:
        if (kind >= ISS_TABLE_MOMENTUM_BOSON0 && kind < ISS_TABLE_MOMENTUM_BOSON0 + 6) {
            // host-provided momentum table [4][n0] -> interleaved [n0][4]
            if (!grid4) return ISS_ERR_ARG;
            const int r = kind - ISS_TABLE_MOMENTUM_BOSON0;
            std::vector<double> inter(static_cast<size_t>(n0)*4);
            for (int64_t i = 0; i < n0; i++)
                for (int c = 0; c < 4; c++) inter[i*4 + c] = data[c*n0 + i];
            int rc = upload_doubles(h, &h->d_momtab[r], inter.data(), inter.size());
            if (rc) return rc;
            MomentumTable &t = h->momtab[r];
            t.data = h->d_momtab[r];
            t.n = static_cast<int>(n0);
            t.m0 = grid4[0];
            t.trunc = static_cast<int>(grid4[1]);
            t.e0 = data[0];
            t.de = data[1] - data[0];
            t.de_build = t.de;
            t.generated = 0;
            t.exp_m0 = exp(t.m0);
            {
                const bool fermion = (r >= 3);
                t.denom0 = fermion ? (1. + exp(-t.m0)) : (1. - exp(-t.m0));
                t.inv_denom0 = 1.0/t.denom0;
                t.inv_de = 1.0/t.de;
                momentum_series_constants(t, fermion);
            }
            return ISS_OK;
        }
        ISS_FAIL(h, ISS_ERR_ARG, "unknown table kind");
    }
}

int iss_cuda_upload_decay_table(iss_handle *h, const iss_decay_species *sp, int32_t nsp,
                                const iss_decay_channel *ch, int32_t nch) {
    if (!h || !sp || !ch || nsp <= 0 || nch <= 0) return ISS_ERR_ARG;
    cudaSetDevice(h->device);
    cudaFree(h->d_dsp); cudaFree(h->d_dch);
    h->d_dsp = nullptr; h->d_dch = nullptr;
    ISS_CUDA_TRY(h, cudaMalloc(&h->d_dsp, sizeof(iss_decay_species)*nsp));
    ISS_CUDA_TRY(h, cudaMalloc(&h->d_dch, sizeof(iss_decay_channel)*nch));
    ISS_CUDA_TRY(h, cudaMemcpyAsync(h->d_dsp, sp, sizeof(iss_decay_species)*nsp,
                                    cudaMemcpyHostToDevice, h->stream));
    ISS_CUDA_TRY(h, cudaMemcpyAsync(h->d_dch, ch, sizeof(iss_decay_channel)*nch,
                                    cudaMemcpyHostToDevice, h->stream));
    // pid-sorted index for look-ups by Monte-Carlo id (the reference searches linearly,
    // particle_decay.cpp:268-273)
    std::vector<std::pair<int32_t, int32_t>> order(nsp);
    for (int i = 0; i < nsp; i++) order[i] = {sp[i].pid, i};
    std::sort(order.begin(), order.end());
    std::vector<int32_t> spid(nsp), sidx(nsp);
    for (int i = 0; i < nsp; i++) {
        spid[i] = order[i].first;
        sidx[i] = order[i].second;
    }
    cudaFree(h->d_sorted_pid); cudaFree(h->d_sorted_idx);
    h->d_sorted_pid = h->d_sorted_idx = nullptr;
    ISS_CUDA_TRY(h, cudaMalloc(&h->d_sorted_pid, sizeof(int32_t)*nsp));
    ISS_CUDA_TRY(h, cudaMalloc(&h->d_sorted_idx, sizeof(int32_t)*nsp));
    ISS_CUDA_TRY(h, cudaMemcpyAsync(h->d_sorted_pid, spid.data(), sizeof(int32_t)*nsp,
                                    cudaMemcpyHostToDevice, h->stream));
    ISS_CUDA_TRY(h, cudaMemcpyAsync(h->d_sorted_idx, sidx.data(), sizeof(int32_t)*nsp,
                                    cudaMemcpyHostToDevice, h->stream));
    ISS_CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    h->ndsp = nsp;
    h->ndch = nch;
    return ISS_OK;
}

int iss_cuda_set_options(iss_handle *h, const iss_options *opt) {
    if (!h || !opt) return ISS_ERR_ARG;
    const int model = opt->dN_dy_sampling_model;
    if (model != 30 && model != 1 && model != 10 && model != 20)
        ISS_FAIL(h, ISS_ERR_ARG,
                 "dN_dy_sampling_model must be 1 (floor+Bernoulli), 10 or 20 (negative binomial) "
                 "or 30 (Poisson) (FSSW.cpp:250-309)");
    const bool changed = !h->have_opt || memcmp(&h->opt, opt, sizeof(*opt)) != 0;
    h->opt = *opt;
    h->have_opt = true;
    h->lambda_on_device = false;
    if (changed) {
        h->have_yields = false;
        h->have_local_yields = false;
        // K/E tables depend on include_deltaf_diffusion: rebuild lazily
    }
    return ISS_OK;
}

int iss_cuda_compute_yields(iss_handle *h, double *dN_species_host, double *yields_host) {
    if (!h) return ISS_ERR_ARG;
    cudaSetDevice(h->device);
    int rc = run_yields(h);
    if (rc) return rc;
    if (dN_species_host)
        memcpy(dN_species_host, h->h_total.data(), sizeof(double)*h->nspecies);
    if (yields_host) {
        ISS_CUDA_TRY(h, cudaMemcpy2DAsync(yields_host, sizeof(double)*h->ncell, h->d_yields,
                                          sizeof(double)*h->ncell_pad, sizeof(double)*h->ncell,
                                          h->nspecies, cudaMemcpyDeviceToHost, h->stream));
        ISS_CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    }
    return ISS_OK;
}

int iss_cuda_set_surface_chunk(iss_handle *h, int64_t cell_begin, int64_t ncell_global) {
    if (!h) return ISS_ERR_ARG;
    if (h->ncell <= 0 || !h->d_surf) ISS_FAIL(h, ISS_ERR_STATE, "upload the cells of the chunk first");
    h->have_yields = false;
    h->have_local_yields = false;
    h->have_batch = false;
    if (ncell_global <= 0) {
        h->chunk = false;
        return ISS_OK;
    }
    if (cell_begin < 0 || cell_begin % ISS_CHUNK_ALIGN != 0)
        ISS_FAIL(h, ISS_ERR_ARG, "cell_begin must be a multiple of ISS_CHUNK_ALIGN (4096)");
    const int64_t cell_end = cell_begin + h->ncell;
    if (cell_end > ncell_global || (cell_end < ncell_global && cell_end % ISS_CHUNK_ALIGN != 0))
        ISS_FAIL(h, ISS_ERR_ARG, "a chunk must end at a multiple of ISS_CHUNK_ALIGN or at the end of the surface");
    if (ncell_global >= (int64_t(1) << 31)) ISS_FAIL(h, ISS_ERR_ARG, "ncell must be < 2^31");
    if (ncell_global <= ISS_CHUNK_ALIGN)
        ISS_FAIL(h, ISS_ERR_ARG, "surface-chunk mode needs a surface of more than 4096 cells");
    h->chunk = true;
    h->chunk_cell_begin = cell_begin;
    h->chunk_tile_begin = cell_begin/TILE;
    h->g_ncell = ncell_global;
    h->g_ntile = (ncell_global + TILE - 1)/TILE;
    return ISS_OK;
}

int iss_cuda_chunk_yields_local(iss_handle *h, void **tilesum_dev, int64_t *ntile_local) {
    if (!h) return ISS_ERR_ARG;
    if (!h->chunk) ISS_FAIL(h, ISS_ERR_STATE, "iss_cuda_set_surface_chunk must run first");
    cudaSetDevice(h->device);
    int rc = run_yields_local(h);
    if (rc) return rc;
    ISS_CUDA_TRY(h, cudaStreamSynchronize(h->stream));   // the caller's collective may use another stream
    if (tilesum_dev) *tilesum_dev = h->d_tilesum;
    if (ntile_local) *ntile_local = h->ntile;
    return ISS_OK;
}

int iss_cuda_chunk_yields_finish(iss_handle *h, const double *const *rank_tilesums,
                                 const int64_t *rank_ntile, int32_t nranks, int on_device,
                                 double *dN_species_host) {
    if (!h || !rank_tilesums || !rank_ntile || nranks <= 0) return ISS_ERR_ARG;
    if (!h->chunk) ISS_FAIL(h, ISS_ERR_STATE, "iss_cuda_set_surface_chunk must run first");
    if (!h->have_local_yields) ISS_FAIL(h, ISS_ERR_STATE, "iss_cuda_chunk_yields_local must run first");
    cudaSetDevice(h->device);
    return iss::chunk_combine_tile_sums(h, rank_tilesums, rank_ntile, nranks, on_device, dN_species_host);
}

int iss_cuda_chunk_block_yields(iss_handle *h, double *block_yield_host, int64_t nblock) {
    if (!h || !block_yield_host) return ISS_ERR_ARG;
    if (!h->chunk || !h->have_yields || h->g_nlev < 3 || !h->d_cdflev_g)
        ISS_FAIL(h, ISS_ERR_STATE, "surface-chunk yields must be complete (iss_cuda_chunk_yields_finish)");
    const int64_t nb = (h->g_ncell + ISS_CHUNK_ALIGN - 1)/ISS_CHUNK_ALIGN;
    if (nblock != nb) ISS_FAIL(h, ISS_ERR_ARG, "nblock must be ceil(ncell_global/4096)");
    cudaSetDevice(h->device);
    // level 3 of the global search tree: inclusive prefix at the end of every block, per species
    const int64_t ns = h->nspecies;
    std::vector<double> lev(static_cast<size_t>(ns)*nb);
    ISS_CUDA_TRY(h, cudaMemcpy2DAsync(lev.data(), sizeof(double)*nb, h->d_cdflev_g + h->g_lev_off[3],
                                      sizeof(double)*h->g_lev_stride, sizeof(double)*nb, ns,
                                      cudaMemcpyDeviceToHost, h->stream));
    ISS_CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    for (int64_t j = 0; j < nb; j++) block_yield_host[j] = 0.;
    for (int64_t s = 0; s < ns; s++) {
        const double *row = lev.data() + s*nb;
        for (int64_t j = 0; j < nb; j++) block_yield_host[j] += row[j] - (j ? row[j - 1] : 0.);
    }
    return ISS_OK;
}

int iss_cuda_sample(iss_handle *h, uint64_t seed, int64_t ev_begin, int64_t ev_end,
                    iss_counts *out) {
    if (!h || ev_end <= ev_begin) return ISS_ERR_ARG;
    if (!h->have_yields) ISS_FAIL(h, ISS_ERR_STATE, "iss_cuda_compute_yields must run first");
    if (ev_end > (int64_t(1) << 32)) ISS_FAIL(h, ISS_ERR_ARG, "event index must be < 2^32");
    cudaSetDevice(h->device);
    h->ev_begin = ev_begin;
    h->ev_end = ev_end;
    h->have_batch = false;
    h->decayed = false;
    const int64_t nev = ev_end - ev_begin;
    int rc = run_multiplicities(h, seed, nev);
    if (rc) return rc;
    rc = run_sampler(h, seed, nev, 0);
    if (rc) return rc;
    unsigned long long cnt[8] = {0};
    rc = mail_post(h, h->d_counters, 8, 0);
    if (rc) return rc;
    ISS_CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    for (int i = 0; i < 8; i++) cnt[i] = *reinterpret_cast<volatile unsigned long long *>(h->h_mail + i);
    h->have_batch = true;
    if (out) {
        out->n_events = nev;
        out->n_hadrons = h->n_hadrons;
        out->n_tries = static_cast<int64_t>(cnt[1]);
        out->n_cell_redraws = static_cast<int64_t>(cnt[2]);
    }
    if (cnt[7] != 0) {
        char buf[240];
        snprintf(buf, sizeof(buf),
                 "surface-chunk mode: %llu hadrons needed a cell re-draw (4999 rejected tries, "
                 "FSSW.cpp:1017-1018) that left this rank's chunk; their records are null (pid 0)", cnt[7]);
        ISS_FAIL(h, ISS_ERR_RANGE, buf);
    }
    if (cnt[6] != 0) {
        // which (cell, species) could not be sampled: the first one a warp gave up on
        unsigned long long who[2] = {0, 0};
        if (mail_post(h, h->d_counters + 8, 2, 24) == ISS_OK
            && cudaStreamSynchronize(h->stream) == cudaSuccess) {
            who[0] = *reinterpret_cast<volatile unsigned long long *>(h->h_mail + 24);
            who[1] = *reinterpret_cast<volatile unsigned long long *>(h->h_mail + 25);
        }
        const int sidx = static_cast<int>(who[1]);
        const int pid = (sidx >= 0 && sidx < h->nspecies) ? h->h_species[sidx].pid : 0;
        char buf[320];
        snprintf(buf, sizeof(buf),
                 "sampler gave up on %llu hadrons after %d rejected tries each (zero acceptance in "
                 "every cell drawn), e.g. species %d (pid %d) in cell %llu; their records are null (pid 0)",
                 cnt[6], 2000000, sidx, pid, who[0]);
        ISS_FAIL(h, ISS_ERR_RANGE, buf);
    }
    if (cnt[3] != 0) {
        char buf[160];
        snprintf(buf, sizeof(buf),
                 "[MomentumSampler] out of range for %llu hadrons (m/T - mu/T outside the tables)",
                 cnt[3]);
        ISS_FAIL(h, ISS_ERR_RANGE, buf);
    }
    return ISS_OK;
}

int iss_cuda_get_multiplicities(iss_handle *h, int64_t *counts_host) {
    if (!h || !counts_host) return ISS_ERR_ARG;
    if (!h->have_batch) ISS_FAIL(h, ISS_ERR_STATE, "no sampled batch");
    const int64_t n = (h->ev_end - h->ev_begin)*h->nspecies;
    ISS_CUDA_TRY(h, cudaMemcpyAsync(counts_host, h->d_mult, sizeof(int64_t)*n,
                                    cudaMemcpyDeviceToHost, h->stream));
    ISS_CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return ISS_OK;
}

int iss_cuda_get_poisson_params(iss_handle *h, double *lambda_host, double *pmode_host) {
    if (!h) return ISS_ERR_ARG;
    if (h->h_lambda.empty()) ISS_FAIL(h, ISS_ERR_STATE, "no sampled batch");
    if (lambda_host) memcpy(lambda_host, h->h_lambda.data(), sizeof(double)*h->nspecies);
    if (pmode_host) memcpy(pmode_host, h->h_pmode.data(), sizeof(double)*h->nspecies);
    return ISS_OK;
}

int iss_cuda_sample_momentum(iss_handle *h, double mass, double T, double mu, int32_t sign,
                             int64_t n, uint64_t seed, double *p_host) {
    if (!h || !p_host || n <= 0) return ISS_ERR_ARG;
    cudaSetDevice(h->device);
    return run_momentum_unit(h, mass, T, mu, sign, n, seed, p_host);
}

int iss_cuda_set_trace(iss_handle *h, int enable) {
    if (!h) return ISS_ERR_ARG;
    h->trace = (enable != 0);
    return ISS_OK;
}

int iss_cuda_get_trace(iss_handle *h, int32_t *cell_host, int32_t *tries_host) {
    if (!h || !cell_host || !tries_host) return ISS_ERR_ARG;
    if (!h->have_batch || !h->trace || !h->d_trace)
        ISS_FAIL(h, ISS_ERR_STATE, "no traced batch (iss_cuda_set_trace before iss_cuda_sample)");
    const size_t nb = sizeof(int32_t)*h->n_primaries;
    ISS_CUDA_TRY(h, cudaMemcpyAsync(cell_host, h->d_trace, nb, cudaMemcpyDeviceToHost, h->stream));
    ISS_CUDA_TRY(h, cudaMemcpyAsync(tries_host, h->d_trace + h->trace_cap, nb,
                                    cudaMemcpyDeviceToHost, h->stream));
    ISS_CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    if (h->chunk)       // cells are reported with their index in the whole surface
        for (int64_t i = 0; i < h->n_primaries; i++)
            cell_host[i] += static_cast<int32_t>(h->chunk_cell_begin);
    return ISS_OK;
}

int iss_cuda_decay(iss_handle *h, uint64_t seed, iss_counts *out) {
    if (!h) return ISS_ERR_ARG;
    if (!h->have_batch) ISS_FAIL(h, ISS_ERR_STATE, "no sampled batch");
    if (!h->d_dsp) ISS_FAIL(h, ISS_ERR_STATE, "decay table not uploaded");
    cudaSetDevice(h->device);
    int rc = run_decay(h, seed);
    if (rc) return rc;
    if (out) {
        out->n_events = h->ev_end - h->ev_begin;
        out->n_hadrons = h->n_hadrons;
        out->n_tries = 0;
        out->n_cell_redraws = 0;
    }
    return ISS_OK;
}

int iss_cuda_event_offsets(iss_handle *h, int64_t *event_offsets_host) {
    if (!h || !event_offsets_host) return ISS_ERR_ARG;
    if (!h->have_batch) ISS_FAIL(h, ISS_ERR_STATE, "no sampled batch");
    const int64_t nev = h->ev_end - h->ev_begin;
    // the kernels that produce the offsets also wrote them to mapped pinned memory, and the
    // batch was synchronised when iss_cuda_sample / iss_cuda_decay returned
    memcpy(event_offsets_host, h->h_evoff, sizeof(int64_t)*(nev + 1));
    return ISS_OK;
}

static const iss_hadron *batch_ptr(const iss_handle *h) {
    return h->decayed ? h->d_hadrons2 : h->d_hadrons;
}

int iss_cuda_fetch_event(iss_handle *h, int64_t iev, iss_hadron *dst, int64_t cap, int64_t *n) {
    if (!h || !n) return ISS_ERR_ARG;
    if (!h->have_batch) ISS_FAIL(h, ISS_ERR_STATE, "no sampled batch");
    const int64_t nev = h->ev_end - h->ev_begin;
    if (iev < 0 || iev >= nev) ISS_FAIL(h, ISS_ERR_ARG, "event index out of range");
    // offsets were mirrored into mapped pinned memory when the batch was produced
    const int64_t off[2] = {h->h_evoff[iev], h->h_evoff[iev + 1]};
    *n = off[1] - off[0];
    if (!dst) return ISS_OK;
    if (*n > cap) ISS_FAIL(h, ISS_ERR_ARG, "destination too small");
    ISS_CUDA_TRY(h, cudaMemcpyAsync(dst, batch_ptr(h) + off[0], sizeof(iss_hadron)*(*n),
                                    cudaMemcpyDeviceToHost, h->stream));
    ISS_CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return ISS_OK;
}

int iss_cuda_fetch_all(iss_handle *h, iss_hadron *dst, int64_t cap, int64_t *n) {
    if (!h || !n) return ISS_ERR_ARG;
    if (!h->have_batch) ISS_FAIL(h, ISS_ERR_STATE, "no sampled batch");
    *n = h->n_hadrons;
    if (!dst) return ISS_OK;
    if (*n > cap) ISS_FAIL(h, ISS_ERR_ARG, "destination too small");
    ISS_CUDA_TRY(h, cudaMemcpyAsync(dst, batch_ptr(h), sizeof(iss_hadron)*(*n),
                                    cudaMemcpyDeviceToHost, h->stream));
    ISS_CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return ISS_OK;
}

int iss_cuda_fetch_all_async(iss_handle *h, iss_hadron *dst, int64_t cap, int64_t *n) {
    if (!h || !n || !dst) return ISS_ERR_ARG;
    if (!h->have_batch) ISS_FAIL(h, ISS_ERR_STATE, "no sampled batch");
    *n = h->n_hadrons;
    if (*n > cap) ISS_FAIL(h, ISS_ERR_ARG, "destination too small");
    ISS_CUDA_TRY(h, cudaEventRecord(h->batch_ready, h->stream));
    ISS_CUDA_TRY(h, cudaStreamWaitEvent(h->copy_stream, h->batch_ready, 0));
    if (*n > 0)
        ISS_CUDA_TRY(h, cudaMemcpyAsync(dst, batch_ptr(h), sizeof(iss_hadron)*(*n),
                                        cudaMemcpyDeviceToHost, h->copy_stream));
    if (h->decayed) {
        ISS_CUDA_TRY(h, cudaEventRecord(h->copy_done2, h->copy_stream));
        h->copy_pending2 = true;
    }
    // the primaries' buffer stays untouched until its copy (or the copy of the decayed batch made
    // from it) is done; the next iss_cuda_sample writes into the other buffer
    ISS_CUDA_TRY(h, cudaEventRecord(h->copy_done[h->cur_buf], h->copy_stream));
    h->copy_pending[h->cur_buf] = true;
    h->cur_buf ^= 1;
    return ISS_OK;
}

int iss_cuda_fetch_wait(iss_handle *h) {
    if (!h) return ISS_ERR_ARG;
    ISS_CUDA_TRY(h, cudaStreamSynchronize(h->copy_stream));
    return ISS_OK;
}

int iss_cuda_device_hadrons(iss_handle *h, const void **dptr, int64_t *n) {
    if (!h || !dptr || !n) return ISS_ERR_ARG;
    if (!h->have_batch) ISS_FAIL(h, ISS_ERR_STATE, "no sampled batch");
    *dptr = batch_ptr(h);
    *n = h->n_hadrons;
    return ISS_OK;
}

int64_t iss_cuda_qa_size(void) { return ISS_QA_HEAD + static_cast<int64_t>(ISS_QA_NSPEC)*ISS_QA_PER; }

int iss_cuda_histograms(iss_handle *h, const int32_t *pids, int32_t npid, int accumulate) {
    if (!h || !pids || npid <= 0 || npid > ISS_QA_NSPEC) return ISS_ERR_ARG;
    if (!h->have_batch) ISS_FAIL(h, ISS_ERR_STATE, "no sampled batch");
    cudaSetDevice(h->device);
    return run_qa(h, pids, npid, accumulate);
}

int iss_cuda_qa_device_ptr(iss_handle *h, void **dptr) {
    if (!h || !dptr) return ISS_ERR_ARG;
    if (!h->d_qa) ISS_FAIL(h, ISS_ERR_STATE, "no QA block (call iss_cuda_histograms)");
    *dptr = h->d_qa;
    return ISS_OK;
}

int iss_cuda_qa_fetch(iss_handle *h, double *dst_host) {
    if (!h || !dst_host) return ISS_ERR_ARG;
    if (!h->d_qa) ISS_FAIL(h, ISS_ERR_STATE, "no QA block (call iss_cuda_histograms)");
    ISS_CUDA_TRY(h, cudaMemcpyAsync(dst_host, h->d_qa, sizeof(double)*iss_cuda_qa_size(),
                                    cudaMemcpyDeviceToHost, h->stream));
    ISS_CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return ISS_OK;
}

int iss_cuda_timing(iss_handle *h, int enable, double *ms_host, int64_t *launches_host,
                    int reset) {
    if (!h) return ISS_ERR_ARG;
    cudaSetDevice(h->device);
    if (!h->spans.empty()) {
        ISS_CUDA_TRY(h, cudaStreamSynchronize(h->stream));
        for (auto &sp : h->spans) {
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, sp.a, sp.b) == cudaSuccess) h->t_ms[sp.kind] += ms;
            h->ev_pool.push_back(sp.a);
            h->ev_pool.push_back(sp.b);
        }
        h->spans.clear();
    }
    if (ms_host) memcpy(ms_host, h->t_ms, sizeof(h->t_ms));
    if (launches_host) memcpy(launches_host, h->t_launch, sizeof(h->t_launch));
    if (reset) {
        memset(h->t_ms, 0, sizeof(h->t_ms));
        memset(h->t_launch, 0, sizeof(h->t_launch));
    }
    h->timing = (enable != 0);
    return ISS_OK;
}

int iss_cuda_mem_info(iss_handle *h, int64_t *free_bytes, int64_t *total_bytes) {
    if (!h) return ISS_ERR_ARG;
    cudaSetDevice(h->device);
    size_t f = 0, t = 0;
    ISS_CUDA_TRY(h, cudaMemGetInfo(&f, &t));
    if (free_bytes) *free_bytes = static_cast<int64_t>(f);
    if (total_bytes) *total_bytes = static_cast<int64_t>(t);
    return ISS_OK;
}

int iss_cuda_host_alloc(iss_handle *h, void **ptr, int64_t bytes) {
    if (!h || !ptr || bytes <= 0) return ISS_ERR_ARG;
    cudaSetDevice(h->device);
    *ptr = nullptr;
    ISS_CUDA_TRY(h, cudaHostAlloc(ptr, static_cast<size_t>(bytes), cudaHostAllocDefault));
    return ISS_OK;
}

int iss_cuda_host_free(iss_handle *h, void *ptr) {
    if (!h) return ISS_ERR_ARG;
    if (ptr) ISS_CUDA_TRY(h, cudaFreeHost(ptr));
    return ISS_OK;
}

int iss_cuda_fp64_peak(iss_handle *h, double *tflops) {
    if (!h || !tflops) return ISS_ERR_ARG;
    cudaSetDevice(h->device);
    int nsm = 148;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, h->device);
    const int blocks = nsm*8, threads = 256, iters = 1 << 16;
    double *d_out = nullptr;
    ISS_CUDA_TRY(h, cudaMalloc(&d_out, sizeof(double)*blocks*threads));
    fp64_fma_kernel<<<blocks, threads, 0, h->stream>>>(d_out, 1024); ISS_LAUNCHED(h);   // warm-up
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 5; rep++) {
        cudaEventRecord(e0, h->stream);
        fp64_fma_kernel<<<blocks, threads, 0, h->stream>>>(d_out, iters); ISS_LAUNCHED(h);
        cudaEventRecord(e1, h->stream);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d_out);
    ISS_CUDA_TRY(h, cudaGetLastError());
    const double flops = 2.0*8.0*static_cast<double>(iters)*blocks*threads;
    *tflops = flops/(best*1e-3)/1e12;
    return ISS_OK;
}

// ---- legacy sampler (MC_sampling = 2) -----------------------------------------------------------
int iss_cuda_legacy_upload_positions(iss_handle *h, const float *pos, int64_t ncell) {
    if (!h || !pos || ncell <= 0) return ISS_ERR_ARG;
    cudaSetDevice(h->device);
    ISS_ENSURE(h, h->d_legpos, h->legpos_bytes, sizeof(float4)*ncell);
    ISS_CUDA_TRY(h, cudaMemcpyAsync(h->d_legpos, pos, sizeof(float4)*ncell, cudaMemcpyHostToDevice,
                                    h->stream));
    ISS_CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    h->nlegpos = ncell;
    if (h->legacy) {
        h->have_yields = false;
        h->have_batch = false;
    }
    return ISS_OK;
}

int iss_cuda_legacy_upload_z_table(iss_handle *h, const double *x, const double *y, int32_t n) {
    if (!h || !x || !y || n < 4) return ISS_ERR_ARG;
    cudaSetDevice(h->device);
    int rc = upload_doubles(h, &h->d_zx, x, n);
    if (rc) return rc;
    rc = upload_doubles(h, &h->d_zy, y, n);
    if (rc) return rc;
    h->nz = n;
    return ISS_OK;
}

int iss_cuda_legacy_set_options(iss_handle *h, const iss_legacy_options *opt) {
    if (!h || !opt) return ISS_ERR_ARG;
    h->legopt = *opt;
    h->have_legopt = true;
    if (h->legacy) h->have_yields = false;
    return ISS_OK;
}

int iss_cuda_legacy_compute_yields(iss_handle *h, double *dN_species_host, double *yields_host,
                                   double *maximum_host) {
    if (!h) return ISS_ERR_ARG;
    cudaSetDevice(h->device);
    int rc = run_legacy_yields(h, yields_host, maximum_host);
    if (rc) return rc;
    if (dN_species_host) memcpy(dN_species_host, h->h_total.data(), sizeof(double)*h->nspecies);
    return ISS_OK;
}

}  // extern "C"
