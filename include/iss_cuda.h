/*
 * iss_cuda.h -- C ABI of the B200 (sm_100a) Cooper-Frye particlization engine.
 *
 * This is the drop-in boundary of the hot path.  The reference (chunshen1987/iSS)
 * has no FFI: its sampler `FSSW` is a C++ class driven by `class iSS`
 * (reference src/iSS.cpp:130-165 -> src/FSSW.cpp:344-361).  Each entry point
 * below names the reference code it replaces.  The C++ facade in
 * iss_b200/host/ (`class iSS`, same public API as reference src/iSS.h:16-102)
 * is the only intended caller; tests and bench.py bind the same symbols with
 * ctypes.
 *
 * Conventions: plain C types, opaque handle, one handle per GPU, caller-owned
 * host buffers, `int` status (0 = ok, !=0 = error; text via iss_cuda_last_error).
 * No exceptions cross this boundary.  A handle is not thread-safe: use it from one
 * thread at a time (different handles may be used concurrently).  There is NO CPU fallback: every call
 * fails with ISS_ERR_CUDA when no sm_100-class device is usable.
 */
#ifndef ISS_CUDA_H_
#define ISS_CUDA_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* entry points are the only symbols libiss_cuda.so exports (it is built with
 * -fvisibility=hidden) */
#if defined(__GNUC__)
#define ISS_API __attribute__((visibility("default")))
#else
#define ISS_API
#endif

typedef struct iss_handle iss_handle;

enum {
    ISS_OK = 0,
    ISS_ERR_CUDA = 1,       /* CUDA runtime error / no device                 */
    ISS_ERR_ARG = 2,        /* bad argument                                   */
    ISS_ERR_STATE = 3,      /* call order violated (e.g. sample before yields) */
    ISS_ERR_RANGE = 4,      /* momentum-sampler table range (reference exit(1),
                               MomentumSamplerBase.cpp:35-43)                 */
    ISS_ERR_NOMEM = 5
};

/* Number of float fields per freeze-out cell in the local-rest-frame record
 * (reference `FO_surf_LRF`, src/data_struct.h:66-76), uploaded as
 * structure-of-arrays.  Index meaning: */
enum {
    ISS_F_TAU = 0, ISS_F_X, ISS_F_Y, ISS_F_ETA,
    ISS_F_DA0, ISS_F_DA1, ISS_F_DA2, ISS_F_DA3,      /* da_mu_LRF[0..3] */
    ISS_F_UT, ISS_F_UX, ISS_F_UY, ISS_F_UZ,          /* u_tz[0..3]      */
    ISS_F_E, ISS_F_T, ISS_F_P, ISS_F_NB,
    ISS_F_MUB, ISS_F_MUS, ISS_F_MUQ, ISS_F_BULKPI,
    ISS_F_PIXX, ISS_F_PIXY, ISS_F_PIXZ, ISS_F_PIYY, ISS_F_PIYZ,
    ISS_F_QX, ISS_F_QY, ISS_F_QZ,
    ISS_NFIELD = 28
};

/* One chosen species, in sampling order (mass-sorted, reference
 * FSSW.cpp:115-162; fields of `particle_info`, data_struct.h:27-50). */
typedef struct {
    int32_t pid;        /* Monte-Carlo id (monval)                       */
    int32_t gspin;
    int32_t baryon;
    int32_t strange;
    int32_t charge;
    int32_t sign;       /* -1 boson, +1 fermion, 0 Boltzmann             */
    int32_t decay_idx;  /* row in the decay table (iss_cuda_upload_decay_table) or -1 */
    int32_t reserved;
    double mass;
} iss_species;

/* Options consumed on the FSSW path (reference FSSW.cpp:67-103, 885-894, 950-951). */
typedef struct {
    int32_t hydro_mode;                 /* 2: 3+1D; otherwise dN *= (y_RB - y_LB), y ~ U */
    int32_t include_deltaf_shear;
    int32_t include_deltaf_bulk;
    int32_t include_deltaf_diffusion;
    int32_t bulk_deltaf_kind;           /* 1, 11, 20, 21 active; 0,2,3,4 no-ops as in FSSW */
    int32_t dN_dy_sampling_model;       /* 30 Poisson (default), 1 floor+Bernoulli,
                                           10 / 20 negative binomial (para1)            */
    int32_t local_charge_conservation;
    int32_t reserved;
    double dN_dy_sampling_para1;
    double y_LB, y_RB;
} iss_options;

/* Table kinds for iss_cuda_upload_table.  `dims`/`grid` meaning per kind:
 *  BESSEL_K   : data[n0][3]  = K1,K2,K3 on x = grid[0] + i*grid[1]       (FSSW.cpp:1609-1643)
 *  EXPINT     : data[n0][9]  = E2,E4..E18 on the same x grid               (FSSW.cpp:1635-1641)
 *  CE         : data[n0*n1][5] rows {e, nB, c2hat, zetahat, etahat}        (FSSW.cpp:1303-1338)
 *  MOM22      : data[n0*n1][8]                                             (FSSW.cpp:1341-1376)
 *  MOM14      : data[3][n0(T)][n1(mu)] c0,c1,c2; grid = {T0,dT,mu0,dmu}    (FSSW.cpp:1215-1300)
 *  KAPPA_B    : data[n0(T)][n1(mu)];           grid = {T0,dT,mu0,dmu}      (FSSW.cpp:1546-1568)
 *  BULK14     : data[n0][4] rows {T [1/fm], B0, D0, E0} of
 *               deltaf_tables/BulkDf_Coefficients_Hadrons_s95p-v0-PCE.dat, n1 = 4: the legacy class's
 *               bulk_deltaf_kind 0 (emissionfunction.cpp:298-301, 3633-3647)
 *  MOMENTUM_* : data[4][n0] = Etilde, CDF_0, CDF_1, CDF_2 of one sampler   (Boson/FermionMomentumSampler.cpp)
 *               regime r in {0,1,2}: kind = ISS_TABLE_MOMENTUM_BOSON0 + r etc.
 */
enum {
    ISS_TABLE_BESSEL_K = 1,
    ISS_TABLE_EXPINT = 2,
    ISS_TABLE_CE = 3,
    ISS_TABLE_MOM22 = 4,
    ISS_TABLE_MOM14 = 5,
    ISS_TABLE_KAPPA_B = 6,
    ISS_TABLE_BULK14 = 7,
    ISS_TABLE_MOMENTUM_BOSON0 = 10,     /* +0,+1,+2 : regimes m0 = 0.05, 30, 50 */
    ISS_TABLE_MOMENTUM_FERMION0 = 13    /* +0,+1,+2 : regimes m0 = 0,    30, 50 */
};

/* Decay table (reference particle_decay.cpp:33-172): one row per species of the
 * pdg list incl. generated anti-baryons; channels flattened. */
typedef struct {
    int32_t pid;
    int32_t stable;
    int32_t n_channels;
    int32_t first_channel;      /* index into the channel array */
    int32_t baryon, strange, charge;
    int32_t reserved;
    double mass;
    double width;
} iss_decay_species;

typedef struct {
    int32_t n_part;             /* as listed (may be 1,2,3,4,-2,...) */
    int32_t daughter[5];        /* row indices into the decay-species array, -1 if none/unknown */
    double branching_ratio;
} iss_decay_channel;

/* Output record: identical to the reference `iSS_Hadron` (data_struct.h:79-84), 40 bytes. */
typedef struct {
    int32_t pid;
    float mass;
    float E, px, py, pz;
    float t, x, y, z;
} iss_hadron;

typedef struct {
    int64_t n_events;           /* events in the sampled batch                  */
    int64_t n_hadrons;          /* hadrons currently held for the batch         */
    int64_t n_tries;            /* accept/reject proposals spent (sampler kernel) */
    int64_t n_cell_redraws;     /* reference "impatience" re-picks (FSSW.cpp:1017-1018) */
} iss_counts;

/* QA block filled by iss_cuda_histograms (per rank; every entry is a plain sum, so the
 * caller all-reduces the block across ranks with NCCL).  Layout, in doubles:
 *   [0]        number of events accumulated
 *   [1..4]     sum over events of P^mu = (E,px,py,pz);  [5..8] sum over events of (P^mu)^2
 *   [9..24]    sum over hadrons of p^mu p^nu / p^0   (iSS::construct_Tmunu..., iSS.cpp:296-331)
 *   [25]       number of hadrons;  [26..28] net baryon number, strangeness, electric charge
 *   [29..31]   reserved
 *   then ISS_QA_NSPEC blocks of ISS_QA_PER doubles, one per tracked pid:
 *     pt_count[NPT], pt_sum[NPT], pt_count_sq[NPT]   Histogram(0,5,100) binning, width 5/99
 *                                                     (Histogram.cpp:7-36); _sq = sum over events
 *                                                     of the per-event bin content squared
 *     y_count[NY] on [-5,5), phi_count[NPHI] on [-pi,pi),
 *     v2_num[NV2], v2_den[NV2]  (sum cos 2phi, count) in pT bins on [0,3),
 *     n_total, n_sq (sum over events of N_ev^2)                                        */
#define ISS_QA_NSPEC 16
#define ISS_QA_NPT 100
#define ISS_QA_NY 100
#define ISS_QA_NPHI 64
#define ISS_QA_NV2 20
#define ISS_QA_HEAD 32
#define ISS_QA_PER (3*ISS_QA_NPT + ISS_QA_NY + ISS_QA_NPHI + 2*ISS_QA_NV2 + 2)

/* ---- lifetime ---------------------------------------------------------- */
ISS_API int iss_cuda_create(int device, iss_handle **out);
ISS_API int iss_cuda_destroy(iss_handle *h);
ISS_API const char *iss_cuda_last_error(const iss_handle *h);
/* Run all work of this handle on an existing CUDA stream (cudaStream_t passed as void*). */
ISS_API int iss_cuda_set_stream(iss_handle *h, void *cuda_stream);
ISS_API int iss_cuda_synchronize(iss_handle *h);

/* ---- inputs (replace the std::vector<FO_surf_LRF>/particle tables FSSW's ctor takes,
 *      FSSW.cpp:43-201) -------------------------------------------------- */
ISS_API int iss_cuda_upload_surface(iss_handle *h, const float *const soa[ISS_NFIELD], int64_t ncell);
/* the same records as one array-of-structures block [ncell][ISS_NFIELD] (the memory order of
 * std::vector<FO_surf_LRF> minus its PCE vector): one host->device copy, transposed on the device.
 * `cells` may be pinned memory; it can be reused when the call returns.                  */
ISS_API int iss_cuda_upload_surface_aos(iss_handle *h, const float *cells, int64_t ncell);
/* the same in parts, so that a host that has to assemble the records (e.g. out of a
 * std::vector<FO_surf_LRF>) overlaps its packing with the copies: parts in ascending order,
 * the first with first = 0, every call with the same ncell_total; cells_part points at record
 * `first` and must stay valid (pinned) until the call of the last part returns, which also
 * synchronises.                                                                          */
ISS_API int iss_cuda_upload_surface_aos_part(iss_handle *h, const float *cells_part, int64_t first,
                                             int64_t n, int64_t ncell_total);
ISS_API int iss_cuda_upload_species(iss_handle *h, const iss_species *species, int32_t nspecies);
ISS_API int iss_cuda_upload_table(iss_handle *h, int32_t kind, const double *data,
                          int64_t n0, int64_t n1, const double *grid4);
ISS_API int iss_cuda_upload_decay_table(iss_handle *h, const iss_decay_species *sp, int32_t nsp,
                                const iss_decay_channel *ch, int32_t nch);
ISS_API int iss_cuda_set_options(iss_handle *h, const iss_options *opt);

/* ---- yields: FSSW::calculate_dN_dxtdy_for_one_particle_species for every species
 *      (FSSW.cpp:565-715, 719-848) + RandomVariable1DArray ctor (RandomVariable1DArray.cpp:25-52).
 *      dN_species_host[ns]  <- sum over cells (NOT yet multiplied by y_RB - y_LB);
 *      yields_host (may be NULL) <- [ns][ncell] FP64 per-cell yields.                */
ISS_API int iss_cuda_compute_yields(iss_handle *h, double *dN_species_host, double *yields_host);

/* ---- surface-chunk sharding: a surface too large, or too slow, for one
 *      GPU is cut into contiguous cell ranges, one per rank; every rank samples ALL events but only
 *      the hadrons whose cell (RandomVariable1DArray::rand over the WHOLE surface,
 *      RandomVariable1DArray.cpp:63-67) lies in its range.  The reference has no counterpart (it is
 *      serial); what is kept is its result: the union of the ranks' hadron lists equals, record for
 *      record, the list of one GPU holding the whole surface, because
 *        - the per-species sum over cells (RandomVariable1DArray.cpp:38-50) is evaluated as tile
 *          sums (1024 cells) combined in one fixed order; ranks exchange the tile sums (one
 *          all-gather of [nspecies][ntile] doubles, the only collective) and each evaluates the
 *          same combination, so totals, Poisson draws and prefix values have the same bits;
 *        - the cell search descends the same 16-ary tree: its upper levels (one entry per 4096
 *          cells) are rebuilt on every rank from the gathered tile sums, the lower ones are local.
 *      Call order per rank:  upload_surface(_aos)(cells of the chunk) -> set_surface_chunk ->
 *      chunk_yields_local -> [all-gather] -> chunk_yields_finish -> sample / decay / histograms.
 *      cell_begin must be a multiple of ISS_CHUNK_ALIGN; the chunk ends at a multiple of it or at
 *      the end of the surface.  ncell_global <= 0 switches the mode off.
 *      Limitation: the reference's re-draw of the cell after 4999 rejected tries
 *      (FSSW.cpp:1017-1018) is only honoured when the new cell is in the same chunk; otherwise the
 *      record is null and iss_cuda_sample returns ISS_ERR_RANGE.                              */
#define ISS_CHUNK_ALIGN 4096
ISS_API int iss_cuda_set_surface_chunk(iss_handle *h, int64_t cell_begin, int64_t ncell_global);
/* yields of the local cells; *tilesum_dev <- device pointer of [nspecies][*ntile_local] doubles
 * (valid until the next yields call), ready to be the send buffer of the all-gather              */
ISS_API int iss_cuda_chunk_yields_local(iss_handle *h, void **tilesum_dev, int64_t *ntile_local);
/* rank_tilesums[r]: the block rank r exported, [nspecies][rank_ntile[r]], in rank (= cell) order;
 * device pointers if on_device != 0 (e.g. the output of an NCCL all-gather), else host memory.
 * dN_species_host[ns] <- sums over the WHOLE surface (as iss_cuda_compute_yields).               */
ISS_API int iss_cuda_chunk_yields_finish(iss_handle *h, const double *const *rank_tilesums,
                                         const int64_t *rank_ntile, int32_t nranks, int on_device,
                                         double *dN_species_host);

/* load balance: block_yield_host[j] <- sum over species of the yield of cells [4096 j, 4096 (j+1))
 * of the WHOLE surface (known on every rank after the finish step; nblock = ceil(ncell_global /
 * 4096)); a host that samples many events per surface re-cuts the chunks with it
 * (iss_b200/sharding.py::split_cells_weighted) so that every rank owns the same number of hadrons. */
ISS_API int iss_cuda_chunk_block_yields(iss_handle *h, double *block_yield_host, int64_t nblock);

/* ---- sampling: FSSW::sample_using_dN_dxtdy_4all_particles_conventional (FSSW.cpp:873-1071)
 *      for events [ev_begin, ev_end): multiplicities, offsets, momenta, boost, emit. */
ISS_API int iss_cuda_sample(iss_handle *h, uint64_t seed, int64_t ev_begin, int64_t ev_end,
                    iss_counts *out);
/* multiplicity table of the last batch: counts_host[(ev-ev_begin)*ns + s]              */
ISS_API int iss_cuda_get_multiplicities(iss_handle *h, int64_t *counts_host);
/* Poisson parameters the draws used: lambda[s], mode pmf pm[s] (for bit-exact CPU checks) */
ISS_API int iss_cuda_get_poisson_params(iss_handle *h, double *lambda_host, double *pmode_host);

/* ---- unit-level access to the |p| sampler (MomentumSamplerShell::Sample_a_momentum,
 *      MomentumSamplerShell.cpp:24-48; the reference's Boson/FermionMomentumSampler_IntegratedTests
 *      exercise exactly this call): n momentum magnitudes for fixed (mass, T, mu, sign).       */
ISS_API int iss_cuda_sample_momentum(iss_handle *h, double mass, double T, double mu, int32_t sign,
                                     int64_t n, uint64_t seed, double *p_host);

/* ---- trace (test instrumentation): when enabled, the sampler also records for every output
 *      slot of the batch the cell it was emitted from and the number of accept/reject
 *      proposals it took; cell_host / tries_host receive n_hadrons int32 each (primaries,
 *      before iss_cuda_decay).                                                          */
ISS_API int iss_cuda_set_trace(iss_handle *h, int enable);
ISS_API int iss_cuda_get_trace(iss_handle *h, int32_t *cell_host, int32_t *tries_host);

/* ---- decays: FSSW::perform_resonance_feed_down + particle_decay (FSSW.cpp:1746-1779,
 *      particle_decay.cpp:265-546) on the batch held by the handle.                    */
ISS_API int iss_cuda_decay(iss_handle *h, uint64_t seed, iss_counts *out);

/* ---- outputs (replace Hadron_list accessors, FSSW.h:166-184) ------------------- */
/* event_offsets_host[n_events+1]: exclusive prefix of hadrons per event of the batch.  */
ISS_API int iss_cuda_event_offsets(iss_handle *h, int64_t *event_offsets_host);
ISS_API int iss_cuda_fetch_event(iss_handle *h, int64_t iev_in_batch, iss_hadron *dst, int64_t cap,
                         int64_t *n);
/* whole batch, event-major; dst may be pinned memory. */
ISS_API int iss_cuda_fetch_all(iss_handle *h, iss_hadron *dst, int64_t cap, int64_t *n);
/* same, asynchronous on the handle's copy stream: returns once the copy is queued behind the
 * batch; the next iss_cuda_sample may run while it is in flight (the library alternates between
 * two device output buffers).  dst must stay valid (and should be pinned) until
 * iss_cuda_fetch_wait returns.                                                            */
ISS_API int iss_cuda_fetch_all_async(iss_handle *h, iss_hadron *dst, int64_t cap, int64_t *n);
ISS_API int iss_cuda_fetch_wait(iss_handle *h);
/* device pointer of the batch (for callers that keep the list on the GPU). */
ISS_API int iss_cuda_device_hadrons(iss_handle *h, const void **dptr, int64_t *n);

/* ---- QA: iSS::perform_checks + Histogram (iSS.cpp:59-83, 296-363): fills a block of
 *      doubles on the DEVICE (so that NCCL can reduce it in place) and optionally copies it. */
ISS_API int64_t iss_cuda_qa_size(void);                   /* number of doubles in the QA block */
ISS_API int iss_cuda_histograms(iss_handle *h, const int32_t *pids, int32_t npid, int accumulate);
ISS_API int iss_cuda_qa_device_ptr(iss_handle *h, void **dptr);
ISS_API int iss_cuda_qa_fetch(iss_handle *h, double *dst_host);

/* ---- the one collective of the path: with one process per GPU and the
 *      events (or surface chunks) sharded over the ranks, iSS::perform_checks (iSS.cpp:59-83)
 *      needs the QA block summed over the ranks; everything else is rank-local.
 *      iss_cuda_histograms_allreduce sums the device block in place over `nccl_comm` (an
 *      ncclComm_t passed as void*) on the handle's stream; a NULL communicator selects the one
 *      made by iss_cuda_nccl_init, and with neither the call is a no-op (single rank).  NCCL is
 *      bound at run time (libnccl.so.2): hosts that never pass more than one rank need none.
 *      iss_cuda_nccl_unique_id / iss_cuda_nccl_init wrap ncclGetUniqueId / ncclCommInitRank for
 *      hosts without NCCL code of their own: rank 0 makes the 128-byte id, the host distributes
 *      it (MPI_Bcast, a file, ...), every rank calls init with it.                           */
ISS_API int iss_cuda_nccl_unique_id(void *id128 /* 128 bytes out */);
ISS_API int iss_cuda_nccl_init(iss_handle *h, const void *id128, int32_t rank, int32_t nranks);
ISS_API int iss_cuda_nccl_finalize(iss_handle *h);
ISS_API int iss_cuda_histograms_allreduce(iss_handle *h, void *nccl_comm);
/* surface-chunk mode, the three steps chunk_yields_local -> all-gather -> chunk_yields_finish as
 * one call on the handle's stream (ncclAllGather of equal [nspecies][max ntile] blocks, no host
 * synchronisation in between): rank_ntile[r] = tiles (1024 cells) of rank r's chunk, in rank order;
 * nccl_comm as for iss_cuda_histograms_allreduce; nranks = 1 needs no communicator.            */
ISS_API int iss_cuda_chunk_yields_allgather(iss_handle *h, const int64_t *rank_ntile, int32_t nranks,
                                            void *nccl_comm, double *dN_species_host);

/* ---- timing: accumulated device time per kernel family.  While enabled, every family span is
 *      bracketed by a pair of CUDA events recorded on the handle's stream WITHOUT synchronising;
 *      the pairs are resolved (one stream synchronisation) by the next call of this function.
 *      launches_host counts every kernel launch of the library, enabled or not.              */
enum { ISS_T_YIELDS = 0, ISS_T_SCAN, ISS_T_MULT, ISS_T_SAMPLE /* proposal kernel */, ISS_T_DECAY,
       ISS_T_QA, ISS_T_SETUP /* sampler set-up kernel */, ISS_T_NKIND };
ISS_API int iss_cuda_timing(iss_handle *h, int enable, double *ms_host /*[ISS_T_NKIND]*/,
                    int64_t *launches_host /*[ISS_T_NKIND]*/, int reset);

/* ---- memory helpers for the host facade (pinned staging buffers, batch sizing) */
ISS_API int iss_cuda_mem_info(iss_handle *h, int64_t *free_bytes, int64_t *total_bytes);
ISS_API int iss_cuda_host_alloc(iss_handle *h, void **ptr, int64_t bytes);   /* cudaHostAlloc */
ISS_API int iss_cuda_host_free(iss_handle *h, void *ptr);

/* ---- FP64 pipe probe: dependent-free DFMA loop, returns achieved TFLOP/s (2 flop per FMA). */
ISS_API int iss_cuda_fp64_peak(iss_handle *h, double *tflops);

/* ---- smooth Cooper-Frye spectra -------------------------
 * Replaces EmissionFunctionArray::calculate_dN_pTdpTdphidy (reference
 * src/emissionfunction.cpp:624-829): dN/(pT dpT dphi dy) on a (pT, phi) grid, summed over all
 * cells and over a table of y - eta_s points, for a list of species.  The cells are the
 * LAB-frame (Milne) records `FO_surf` (src/data_struct.h:53-63) that iSS::read_in_FO_surface
 * keeps when MC_sampling != 4 (src/iSS.cpp:105-109), one array-of-structures block of
 * ISS_LAB_NFIELD floats per cell in this order: */
enum {
    ISS_L_TAU = 0, ISS_L_U0, ISS_L_U1, ISS_L_U2, ISS_L_U3,
    ISS_L_DA0, ISS_L_DA1, ISS_L_DA2, ISS_L_DA3,
    ISS_L_T, ISS_L_P, ISS_L_E, ISS_L_MUB, ISS_L_MUS, ISS_L_MUQ,
    ISS_L_PI00, ISS_L_PI01, ISS_L_PI02, ISS_L_PI03, ISS_L_PI11, ISS_L_PI12, ISS_L_PI13,
    ISS_L_PI22, ISS_L_PI23, ISS_L_PI33,
    ISS_L_BULKPI, ISS_L_BN, ISS_L_Q0, ISS_L_Q1, ISS_L_Q2, ISS_L_Q3, ISS_L_SPARE,
    ISS_LAB_NFIELD = 32
};

/* Parameters read by the legacy class (src/emissionfunction.cpp:88-100, 632). */
typedef struct {
    int32_t include_deltaf_shear;
    int32_t include_deltaf_bulk;
    int32_t bulk_deltaf_kind;       /* 1..4 polynomial coefficients; 0: coefficients stay 0 (the
                                       reference never fills them on this path); others: no bulk */
    int32_t include_deltaf_diffusion;   /* needs ISS_TABLE_KAPPA_B */
    int32_t restrict_deltaf;
    int32_t use_pos_dN_only;
    double deltaf_max_ratio;
} iss_spectra_options;

ISS_API int iss_cuda_upload_surface_lab(iss_handle *h, const float *cells, int64_t ncell);
/* dN and dN_max: host arrays [nspecies][npT][nphi] (dN_max may be NULL).  y_minus_eta / y_weight
 * are the two columns of the reference's eta table (bin_tables/eta_uni_table.dat).  Uses
 * iss_species.{mass, gspin, baryon, strange, charge, sign}.  The sum over cells runs in chunks
 * whose size depends on ncell only, so the result is bit-identical from run to run and
 * independent of how species are spread over GPUs. */
ISS_API int iss_cuda_spectra(iss_handle *h, const iss_spectra_options *opt,
                             const iss_species *species, int32_t nspecies,
                             const double *pT, int32_t npT, const double *phi, int32_t nphi,
                             const double *y_minus_eta, const double *y_weight, int32_t ny,
                             double *dN, double *dN_max);
/* evaluations (cell x y-eta point x pT x phi x species) and kernel milliseconds of the last call */
ISS_API int iss_cuda_spectra_stats(iss_handle *h, double *evaluations, double *kernel_ms);

/* ---- legacy "conventional" sampler --------------------------
 * Replaces EmissionFunctionArray::sample_using_dN_dxtdy_4all_particles_conventional, the
 * MC_sampling = 2 path (reference src/emissionfunction.cpp:3273-3623): yields over the LAB-frame
 * (Milne) cells (calculate_dN_dxtdy_for_one_particle_species, :2977-3208), cell choice
 * (RandomVariable1DArray), estimate_maximum (:4006-4153, 4309-4421), uniform proposals in
 * (pT^2, phi, y - eta_s) accepted against that maximum (sample_momemtum_from_a_fluid_cell,
 * :4188-4306), add_one_sampled_particle (:4423-4475).
 * Call order: upload_species / upload_table(BESSEL_K [, EXPINT, KAPPA_B]) / set_options (hydro_mode,
 * y_LB, y_RB, multiplicity model, local_charge_conservation) as for the FSSW path, then
 *   iss_cuda_upload_surface_lab(cells) -> iss_cuda_legacy_upload_positions ->
 *   iss_cuda_legacy_upload_z_table -> iss_cuda_legacy_set_options -> iss_cuda_legacy_compute_yields
 *   -> iss_cuda_sample / iss_cuda_decay / iss_cuda_histograms / fetch as usual.
 * bulk_deltaf_kind 0 needs iss_cuda_upload_table(ISS_TABLE_BULK14).  Not supported: PCE chemical
 * potentials (unreachable in the reference as shipped).                                       */
typedef struct {
    int32_t include_deltaf_shear;
    int32_t include_deltaf_bulk;
    int32_t bulk_deltaf_kind;           /* 0..4 (the yields carry a bulk term for kind 1 only,
                                           emissionfunction.cpp:3143-3147; 0: table ISS_TABLE_BULK14);
                                           others: zero coefficients */
    int32_t include_deltaf_diffusion;   /* needs ISS_TABLE_KAPPA_B and ISS_TABLE_EXPINT */
    int32_t restrict_deltaf;
    int32_t reserved;
    double deltaf_max_ratio;
    double sample_pT_up_to;             /* > 0: the host resolves the reference's "-1 = last row of
                                           the pT table" (emissionfunction.cpp:3302-3305) */
    double sample_y_minus_eta_s_range;
} iss_legacy_options;

/* pos: [ncell][4] floats = xpt, ypt, eta_s, 0 of the cells given to iss_cuda_upload_surface_lab */
ISS_API int iss_cuda_legacy_upload_positions(iss_handle *h, const float *pos, int64_t ncell);
/* the two columns of iSS_tables/z_exp_m_z.dat (TableFunction z_exp_m_z, emissionfunction.cpp:3284-3287) */
ISS_API int iss_cuda_legacy_upload_z_table(iss_handle *h, const double *x, const double *y, int32_t n);
ISS_API int iss_cuda_legacy_set_options(iss_handle *h, const iss_legacy_options *opt);
/* dN_species_host[ns] <- sum over cells of max(yield, 0) (RandomVariable1DArray::return_sum);
 * yields_host (may be NULL) <- [ns][ncell] raw yields (the reference does not clamp them);
 * maximum_host (may be NULL) <- [ns][ncell] estimate_maximum values (test instrumentation).     */
ISS_API int iss_cuda_legacy_compute_yields(iss_handle *h, double *dN_species_host,
                                           double *yields_host, double *maximum_host);

/* ---- surface ingest on the device -------------------------
 * Binary MUSIC surface records (34 float32 per cell, reference src/readindata.cpp:646-689) ->
 * local-rest-frame records in ISS_F_* order, i.e. the per-cell work of
 * read_FOdata::read_FOsurfdat_MUSIC(_boost_invariant) (readindata.cpp:395-546, 626-765),
 * regulate_surface_cells / getValuesFromHRGEOS / regulate_Wmunu (:768-842, 1216-1309),
 * iSS::computeFOSurfTmunu (src/iSS.cpp:378-445, per-cell tensors) and
 * iSS::transform_to_local_rest_frame (src/iSS.cpp:170-293), with both filters applied in file
 * order: cells with T <= 0.01 GeV (readindata.cpp:752) and cells with u.dsigma < 0 (iSS.cpp:226). */
typedef struct {
    int32_t boost_invariant;    /* hydro_mode 1: eta_s = 0, da3 = 0 (readindata.cpp:428-432)       */
    int32_t regulate_eos;       /* EOS 9/91/12/14: T, mu_B, mu_S, mu_Q, P from the HRG table        */
    int32_t hrg_nB;             /* n_B points per energy-density row: 1 (EOS 9/91) or 200 (12/14)   */
    int32_t reserved;
    int64_t hrg_rows;           /* rows of 7 doubles: ed, nB, P, T, muB, muS, muQ                   */
} iss_ingest_options;

typedef struct {
    int64_t n_in;               /* records in the file                                             */
    int64_t n_after_T;          /* cells that passed the T filter (tmunu_out rows)                  */
    int64_t n_kept;             /* cells that also passed u.dsigma >= 0 (lrf_out rows)              */
} iss_ingest_result;

/* per-record status bits (status_out) */
#define ISS_INGEST_DROPPED_T 1          /* T <= 0.01 GeV                                           */
#define ISS_INGEST_EOS_RANGE 2          /* energy density outside the HRG table: T, mu, P kept       */
#define ISS_INGEST_DROPPED_NORMAL 4     /* u.dsigma < 0                                             */

/* All pointers are host memory.  lrf_out: room for ncell x ISS_NFIELD floats; tmunu_out (may be
 * NULL): ncell x 16 floats, the per-cell T^{mu nu} of iSS::computeFOSurfTmunu for the cells that
 * passed the T filter; status_out (may be NULL): ncell bytes. */
ISS_API int iss_cuda_ingest_music_binary(iss_handle *h, const float *raw, int64_t ncell,
                                         const iss_ingest_options *opt, const double *hrg,
                                         float *lrf_out, float *tmunu_out, uint8_t *status_out,
                                         iss_ingest_result *res);

#ifdef __cplusplus
}
#endif
#endif  /* ISS_CUDA_H_ */
