// iss_internal.cuh -- handle layout and helpers shared by the CUDA translation units.
#ifndef ISS_INTERNAL_CUH_
#define ISS_INTERNAL_CUH_

#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <cstring>
#include <vector>
#include <cstdio>

#include "../../include/iss_cuda.h"
#include "iss_rng.h"

namespace iss {

constexpr double HBARC = 0.197327053;     // reference data_struct.h:9
constexpr int TILE = 1024;                // cells per CDF tile (scan granularity)
constexpr int MAX_SPECIES = 1024;
constexpr int MAIL_WORDS = 64;          // [0..7] sampler counters, [8] scan total, [16..17] decay errors,
                                        // [24..25] give-up info
constexpr int N_COUNTERS = 16;          // device counters: [0..7] see SamplerArgs, [8..9] give-up info

// special-function table grid (FSSW.cpp:1611-1615)
struct SfGrid {
    double x_min, dx, x_max_minus_dx;
    int n;
};

// 2-D coefficient grid (T, mu) for kappa_B and the 14-moment tables
struct Grid2D {
    double x0, dx, y0, dy;
    int nx, ny;
};

struct MomentumTable {      // one Boson/FermionMomentumSampler instance
    const double *data;     // [n][4]: Etilde, CDF_0, CDF_1, CDF_2 (interleaved, 32 B per point)
    int n;
    int trunc;              // trunc_order_
    double m0;              // regime base (m0_)
    double e0, de;          // Etilde_[0], Etilde_[1] - Etilde_[0]
    // constants of the series evaluated per (cell, species) at Etilde = (m - mu)/T
    double exp_m0;          // exp(m0)
    double denom0;          // 1 -/+ exp(-m0) (boson / fermion closed form of CDF_0)
    double a[10];           // exp(-m0 n), n = 0..9
    double inv_n1[10];      // 1/(n+1)
    double inv_denom0, inv_de;
    double de_build;        // Etilde_[i] = fma(i, de_build, e0) for tables generated on the device
    int generated;          // 0: uploaded by the host (arbitrary abscissa, never copied to smem)
    // the 10-term series as polynomials in exp(m0 - Etilde) (momentum_series_constants):
    // q1[n] = a[n]/(n+1), q2[n] = a[n]/(n+1)^2, q3[n] = a[n]/(n+1)^3, k1 / k2 = the Etilde-independent
    // parts of CDF_1 / CDF_2
    double q1[10], q2[10], q3[10], k1, k2;
};

// fills the series constants of a table from m0, trunc and the statistics
inline void momentum_series_constants(MomentumTable &t, bool fermion) {
    t.k1 = 0.;
    t.k2 = 0.;
    double sign = 1.;
    for (int n = 0; n < 10; n++) {
        t.a[n] = exp(-t.m0*n);
        t.inv_n1[n] = 1.0/(n + 1);
        const double inv = t.inv_n1[n], n1 = n + 1;
        t.q1[n] = t.a[n]*inv;
        t.q2[n] = t.a[n]*inv*inv;
        t.q3[n] = t.a[n]*inv*inv*inv;
        if (n < t.trunc) {
            t.k1 += (fermion ? sign : 1.)*inv*inv*t.a[n]*(t.m0*n1 + 1);
            t.k2 += inv*inv*inv*t.a[n]*(t.m0*n1*(t.m0*n1 + 2) + 2);
        }
        sign = -sign;
    }
}

constexpr int CELL_STRIDE = 32;   // floats per AoS cell record (28 fields + t, z + 2 spare)
constexpr int CELL_T = 28;        // tau*cosh(eta)
constexpr int CELL_Z = 29;        // tau*sinh(eta)
constexpr int COEF_STRIDE = 8;

struct DeviceSpecies {      // what the kernels read per species (smem-friendly, 48 B)
    double mass;
    double mass2;           // mass*mass
    double inv_mass;        // 1/mass
    int32_t pid;
    int16_t gspin, baryon, strange, charge, sign;
    int16_t trunc10_mass;   // 1 if mass < 0.7 (series of 10 terms when also T > 0.05)
    int32_t decay_idx;
    int32_t combo;          // index of (B,S,Q) among the distinct combinations of the list
};

}  // namespace iss

struct iss_handle {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    std::string err;

    // surface
    int64_t ncell = 0;
    float *d_surf = nullptr;            // [ISS_NFIELD][ncell_pad]
    int64_t ncell_pad = 0;              // multiple of TILE
    int64_t ntile = 0;
    float *d_cells = nullptr;           // [ncell][32] AoS copy for the sampler's random access
    double *d_cellcoef = nullptr;       // [ncell][8] delta-f coefficients c0..c5, kappa, spare
    float *d_stage = nullptr;           // [ncell][28] staging of an AoS upload
    float4 *d_thermo = nullptr;         // [ncell] {T, muB, muS, muQ}: 16 B per cell, L2 resident, all the
                                        // sampler's set-up kernel needs from a cell
    size_t surf_bytes = 0, cells_bytes = 0, coef_bytes = 0, stage_bytes = 0, thermo_bytes = 0;   // capacities

    // species
    int nspecies = 0;
    std::vector<iss_species> h_species;
    iss::DeviceSpecies *d_species = nullptr;

    // options
    iss_options opt{};
    bool have_opt = false;

    // tables
    double *d_bessel = nullptr; iss::SfGrid sf{};
    double *d_expint = nullptr;
    double *d_sf4 = nullptr; bool sf4_with_diff = false;
    int4 *d_combos = nullptr; int ncombo = 0;
    double *d_ce = nullptr; int ce_ne = 0, ce_nb = 0;
    double *d_mom22 = nullptr;
    double *d_mom14 = nullptr; iss::Grid2D g14{};
    double *d_kappa = nullptr; iss::Grid2D gk{};
    // smooth spectra (spectra.cu)
    float *d_lab = nullptr; size_t lab_bytes = 0; int64_t nlab = 0;     // [nlab][ISS_LAB_NFIELD]
    // legacy sampler (MC_sampling = 2, legacy.cuh) on the lab-frame surface d_lab
    bool legacy = false;                // the yields / CDF held are those of the legacy path
    int64_t ncell_lrf = 0;              // cells of the uploaded FSSW (local-rest-frame) surface
    iss_legacy_options legopt{}; bool have_legopt = false;
    float4 *d_legpos = nullptr; size_t legpos_bytes = 0; int64_t nlegpos = 0;   // x, y, eta_s, 0
    double4 *d_legcoef = nullptr; size_t legcoef_bytes = 0;                     // c0, c1, c2, kappa
    double *d_zx = nullptr, *d_zy = nullptr; int nz = 0;                        // z_exp_m_z.dat
    double *d_bulk0 = nullptr; int nbulk0 = 0;      // [4][n] bulk coefficients of kind 0 (T, B0, D0, E0)
    double *d_lambert = nullptr;                                                // W0 table
    double *d_legmax = nullptr; size_t legmax_bytes = 0;
    double *d_labrec = nullptr; size_t labrec_bytes = 0;                // per-cell double records
    double *d_spec_part = nullptr; size_t spec_part_bytes = 0;          // chunk partial sums
    double *d_spec_out = nullptr; size_t spec_out_bytes = 0;
    double *d_spec_tab = nullptr; size_t spec_tab_bytes = 0;            // momentum / eta tables
    double spec_evals = 0., spec_ms = 0.;
    void *d_ingest = nullptr; size_t ingest_bytes = 0;                  // surface ingest arena (ingest.cu)
    double *d_momtab[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    iss::MomentumTable momtab[6]{};     // 0..2 boson regimes, 3..5 fermion regimes

    // decay table
    iss_decay_species *d_dsp = nullptr; int ndsp = 0;
    iss_decay_channel *d_dch = nullptr; int ndch = 0;
    int32_t *d_sorted_pid = nullptr;    // pid-sorted view of the decay table
    int32_t *d_sorted_idx = nullptr;

    // yields / CDF
    double *d_yields = nullptr;         // [ns][ncell_pad]  raw yields
    double *d_cdf = nullptr;            // [ns][ncell_pad]  inclusive scan inside each tile
    double *d_tilesum = nullptr;        // [ns][ntile]
    double *d_tilebase = nullptr;       // [ns][ntile+1]    exclusive prefix over tiles
    double *d_total = nullptr;          // [ns]
    // 16-ary search levels over the global inclusive prefix d_cdf: level k (k >= 1) holds every
    // 16^k-th prefix value; all levels of one species are contiguous, padded to multiples of 16
    double *d_cdflev = nullptr; size_t cdflev_bytes = 0;
    int nlev = 0;                       // number of levels above the prefix itself
    int64_t lev_n[8] = {0}, lev_off[8] = {0}, lev_stride = 0;
    size_t yields_bytes = 0, cdf_bytes = 0, tilesum_bytes = 0, tilebase_bytes = 0, total_bytes = 0;
    // guide table of the cell search over level 1 (yields.cu, guide_kernel): [ns][guide_M + 1]
    // pairs {G[k], G[k+1]}; guide_M = 0: not built (surface-chunk mode)
    void *d_guide = nullptr; size_t guide_bytes = 0; int64_t guide_M = 0;
    bool have_yields = false;
    bool have_local_yields = false;     // yields + tile sums of the local cells (part 1 of run_yields)
    // surface-chunk sharding (iss_cuda_set_surface_chunk): this handle holds cells
    // [chunk_cell_begin, chunk_cell_begin + ncell) of a surface of g_ncell cells.  Tile sums,
    // tile bases and search levels >= 3 exist for the WHOLE surface (small), everything per cell
    // only for the local cells.
    bool chunk = false;
    int64_t chunk_cell_begin = 0, chunk_tile_begin = 0, g_ncell = 0, g_ntile = 0;
    double *d_tilesum_g = nullptr;      // [ns][g_ntile]
    double *d_tilebase_g = nullptr;     // [ns][g_ntile+1]
    double *d_cdflev_g = nullptr;       // [ns][g_lev_stride]: levels 3..g_nlev of the global tree
    size_t tilesum_g_bytes = 0, tilebase_g_bytes = 0, cdflev_g_bytes = 0;
    int g_nlev = 0;
    int64_t g_lev_n[8] = {0}, g_lev_off[8] = {0}, g_lev_stride = 0;
    double *d_tilesum_all = nullptr; size_t tilesum_all_bytes = 0;  // all-gather of the ranks' tile sums
    int64_t *d_own = nullptr; int64_t own_cap = 0;        // [ns*nev + 1] owned draws per (species, event) -> prefix
    uint32_t *d_ownmask = nullptr; int64_t ownmask_cap = 0;   // ballot words of the ownership pass
    int64_t *d_wlist = nullptr; int64_t wlist_cap = 0;    // uint4 identity per owned hadron (two int64 each)
    std::vector<double> h_total;        // dN per species (3+1D sum)
    std::vector<double> h_lambda, h_pmode;
    double *d_lambda = nullptr, *d_pmode = nullptr;
    bool lambda_on_device = false;      // reset whenever yields or options change

    // sampling batch
    int64_t ev_begin = 0, ev_end = 0;
    int64_t *d_mult = nullptr;          // [nev][ns]
    int64_t *d_off_out = nullptr;       // [nev*ns + 1] event-major exclusive prefix
    int64_t *d_off_work = nullptr;      // [ns*nev + 1] species-major exclusive prefix
    int64_t mult_cap = 0;
    iss_hadron *d_hadrons = nullptr;    // = d_hadbuf[cur_buf]
    int64_t hadron_cap = 0;
    // two output buffers so that the device->host copy of one batch (copy_stream) overlaps the
    // sampling of the next (iss_cuda_fetch_all_async)
    iss_hadron *d_hadbuf[2] = {nullptr, nullptr};
    int64_t hadbuf_cap[2] = {0, 0};
    int cur_buf = 0;
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t batch_ready = nullptr;          // recorded on stream when a batch is complete
    cudaEvent_t copy_done[2] = {nullptr, nullptr};
    bool copy_pending[2] = {false, false};
    cudaEvent_t copy_done2 = nullptr;           // copies out of d_hadrons2 (decayed batches)
    bool copy_pending2 = false;
    int64_t n_hadrons = 0;
    int64_t *d_event_off = nullptr;     // [nev+1]
    int64_t event_off_cap = 0;
    void *d_sampler_args = nullptr;
    void *d_hints = nullptr; size_t hints_bytes = 0;
    void *d_tasks = nullptr; size_t tasks_bytes = 0;     // cell-sorted task list of the batch (Task32)
    void *d_tasks_unsorted = nullptr;                    // the same tasks in work order (same capacity)
    uint32_t *d_task_slot = nullptr; size_t task_slot_bytes = 0;        // surface-chunk mode: output slots, sorted
    unsigned long long *d_cellcnt = nullptr; size_t cellcnt_bytes = 0;  // [ncell + 2] histogram / offsets
    void *d_cellrec = nullptr; size_t cellrec_bytes = 0; // [ncell] CellRec (sampler.cu)
    bool cellrec_valid = false;                          // reset with the yields
    unsigned long long *d_counters = nullptr;   // [8] tries, redraws, error flags...
    bool have_batch = false;
    bool trace = false;
    int32_t *d_trace = nullptr;         // [2][trace_cap]: cell, tries per output slot
    int64_t trace_cap = 0;
    int64_t n_primaries = 0;
    bool decayed = false;
    iss_hadron *d_hadrons2 = nullptr;   // decay output
    int64_t hadron2_cap = 0;
    int64_t *d_decay_cnt = nullptr;     // per-primary final multiplicity / offsets
    int64_t decay_cnt_cap = 0;
    void *d_scan_tmp = nullptr; size_t scan_tmp_bytes = 0;

    // mailbox: a page of mapped pinned host memory that tiny kernels write scalars into (scan
    // totals, counters), so that no small device->host copy has to queue on the copy engine
    // behind a multi-hundred-MB hadron transfer
    unsigned long long *h_mail = nullptr;       // host view  [ISS_MAIL_WORDS]
    unsigned long long *d_mail = nullptr;       // device view of the same page
    int64_t *h_evoff = nullptr, *d_evoff_mapped = nullptr;   // mapped copy of the event offsets
    int64_t evoff_mapped_cap = 0;

    // QA
    double *d_qa = nullptr;
    double *d_qa_scratch = nullptr; size_t qa_scratch_bytes = 0;   // per-CTA FP64 sums of the QA kernel
    // communicator made by iss_cuda_nccl_init (collective.cu); null: single rank
    void *nccl_comm = nullptr;
    int nccl_rank = 0, nccl_nranks = 0;

    // timing: event pairs are recorded on the stream without synchronising and resolved in
    // iss_cuda_timing(); t_launch counts every kernel launch of the library
    bool timing = false;
    double t_ms[ISS_T_NKIND] = {0};
    int64_t t_launch[ISS_T_NKIND] = {0};
    struct Span { int kind; cudaEvent_t a, b; };
    std::vector<Span> spans;
    std::vector<cudaEvent_t> ev_pool;
    int cur_kind = ISS_T_YIELDS;
};

// every <<< >>> of the library is followed by ISS_LAUNCHED(h): counts the launch in the family
// of the enclosing ScopedTimer
#define ISS_LAUNCHED(h) ((h)->t_launch[(h)->cur_kind]++)

#define ISS_CUDA_TRY(h, expr)                                                        \
    do {                                                                             \
        cudaError_t e__ = (expr);                                                    \
        if (e__ != cudaSuccess) {                                                    \
            (h)->err = std::string(#expr) + ": " + cudaGetErrorString(e__);          \
            return ISS_ERR_CUDA;                                                     \
        }                                                                            \
    } while (0)

#define ISS_FAIL(h, code, msg)                                                       \
    do {                                                                             \
        (h)->err = (msg);                                                            \
        return (code);                                                               \
    } while (0)

namespace iss {

struct ScopedTimer {
    iss_handle *h;
    int kind, prev_kind;
    cudaEvent_t a = nullptr, b = nullptr;
    static cudaEvent_t get(iss_handle *h) {
        cudaEvent_t e = nullptr;
        if (!h->ev_pool.empty()) {
            e = h->ev_pool.back();
            h->ev_pool.pop_back();
        } else {
            cudaEventCreate(&e);
        }
        return e;
    }
    ScopedTimer(iss_handle *h_, int kind_) : h(h_), kind(kind_), prev_kind(h_->cur_kind) {
        h->cur_kind = kind;
        if (h->timing) {
            a = get(h);
            b = get(h);
            cudaEventRecord(a, h->stream);
        }
    }
    ~ScopedTimer() {
        h->cur_kind = prev_kind;
        if (a) {
            cudaEventRecord(b, h->stream);
            h->spans.push_back({kind, a, b});
        }
    }
};

// copies n 64-bit words device -> mailbox[slot..] with a one-warp kernel (no copy engine)
int mail_post(iss_handle *h, const void *d_src, int n, int slot);
int ensure_mapped_event_offsets(iss_handle *h, int64_t n);

// exclusive scan of int64 on the handle's stream (own kernels, see scan.cu)
int device_exclusive_scan_i64(iss_handle *h, const int64_t *d_in, int64_t *d_out, int64_t n,
                              int64_t *h_total);

int run_yields(iss_handle *h);
int run_yields_local(iss_handle *h);
int run_yields_finish(iss_handle *h);
int chunk_combine_tile_sums(iss_handle *h, const double *const *rank_tilesums, const int64_t *rank_ntile,
                            int32_t nranks, int on_device, double *dN_species_host);
int run_multiplicities(iss_handle *h, uint64_t seed, int64_t nev);
int run_sampler(iss_handle *h, uint64_t seed, int64_t nev, int64_t total);
int build_cellrec(iss_handle *h);
int run_legacy_yields(iss_handle *h, double *yields_host, double *maximum_host);
int run_decay(iss_handle *h, uint64_t seed);
int run_qa(iss_handle *h, const int32_t *pids, int npid, int accumulate);
int run_momentum_unit(iss_handle *h, double m, double T, double mu, int sign, int64_t n,
                      uint64_t seed, double *out_host);

// grows *ptr to at least `need` bytes (contents are not preserved)
inline int ensure_bytes(iss_handle *h, void **ptr, size_t *cap, size_t need) {
    if (*ptr && need <= *cap) return ISS_OK;
    if (*ptr) cudaFree(*ptr);
    *ptr = nullptr;
    *cap = 0;
    cudaError_t e = cudaMalloc(ptr, need);
    if (e != cudaSuccess) {
        h->err = std::string("cudaMalloc failed: ") + cudaGetErrorString(e);
        return ISS_ERR_NOMEM;
    }
    *cap = need;
    return ISS_OK;
}
#define ISS_ENSURE(h, ptr, cap, need)                                                    \
    do {                                                                                 \
        int rc__ = iss::ensure_bytes((h), reinterpret_cast<void **>(&(ptr)), &(cap), (need)); \
        if (rc__) return rc__;                                                           \
    } while (0)

template <typename T>
int ensure_capacity(iss_handle *h, T **ptr, int64_t *cap, int64_t need) {
    if (need <= *cap && *ptr) return ISS_OK;
    if (*ptr) cudaFree(*ptr);
    *ptr = nullptr;
    int64_t newcap = need + need/8 + 1024;
    cudaError_t e = cudaMalloc(reinterpret_cast<void **>(ptr), sizeof(T)*newcap);
    if (e != cudaSuccess) {
        *cap = 0;
        h->err = std::string("cudaMalloc failed: ") + cudaGetErrorString(e);
        return ISS_ERR_NOMEM;
    }
    *cap = newcap;
    return ISS_OK;
}

}  // namespace iss
#endif  // ISS_INTERNAL_CUH_
