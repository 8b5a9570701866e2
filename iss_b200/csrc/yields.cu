// yields.cu -- per-cell x per-species thermal yields and their cell CDF.
//
// Replaces FSSW::calculate_dN_dxtdy_for_one_particle_species (FSSW.cpp:565-715),
// FSSW::calculate_dN_analytic (FSSW.cpp:719-848) and the RandomVariable1DArray
// constructor (RandomVariable1DArray.cpp:25-52) for ALL species in one pass:
//   K3a yields_kernel   : one thread per cell, loop over a chunk of species; the
//                         species-independent delta-f coefficients are computed
//                         once per cell (the reference recomputes them per species).
//   K3b tile_scan_kernel: inclusive scan of the yields inside tiles of 1024 cells
//                         (warp-shuffle scan) + per-tile sums.
//   K3c tile_base_kernel: one warp per species, fixed-order exclusive prefix over
//                         the tile sums -> tile bases and species totals.
// All reductions have a fixed order, so totals are bit-reproducible run to run
// and identical on every GPU that holds the same surface.
#include "coefficients.cuh"
#include "legacy.cuh"

namespace iss {

struct YieldArgs {
    const float *surf;      // [ISS_NFIELD][ncell_pad]
    int64_t ncell, ncell_pad;
    const DeviceSpecies *species;
    int ns;
    int species_per_block;
    CoefTables tab;
    ModeFlags mode;
    double *yields;         // [ns][ncell_pad]
    double *cellcoef;       // [ncell][COEF_STRIDE] by-product for the sampler
    const double *sf4;      // [n][4] K1, K2, K3, weighted sum of E_2..E_18
    const int4 *combos;     // [ncombo] distinct (B, S, Q) of the species list
    int ncombo;
};

// 1/n for the terms of the quantum-statistics series (same bits as the division 1.0/n)
__constant__ double c_inv_n[11] = {0., 1.0, 1.0/2, 1.0/3, 1.0/4, 1.0/5, 1.0/6, 1.0/7, 1.0/8, 1.0/9, 1.0/10};
constexpr int YIELD_THREADS = 128;
constexpr int YIELD_SPECIES_SMEM = 512;

// I_1 weights of E_2, E_4, ..., E_18 (FSSW.cpp:787-806): 3/8, then 3 (2k-5)!!/(2^k k!), k = 3..10
__host__ __device__ inline void expint_weights(double w[9]) {
    w[0] = 3./8.;
    double double_factorial = 1., factorial = 2., two_k = 4.;
    for (int k = 3; k <= 10; k++) {
        double_factorial *= (2*k - 5);
        factorial *= k;
        two_k *= 2;
        w[k - 2] = 3.*double_factorial/two_k/factorial;
    }
}

// One thread per cell, species loop.  Per cell (hoisted out of the species loop, the reference
// recomputes them per species): delta-f coefficients and the fugacities exp(mu/T) of the distinct
// (B,S,Q) combinations of the list (~30 instead of one exp per species), kept in shared memory.
// Per term of the series one pair of 32-byte table records {K1,K2,K3,I} is read: the nine E_2k
// look-ups of the diffusion term are one look-up of their pre-combined weighted sum (a lerp is
// linear, so lerp-then-combine == combine-then-lerp up to rounding).
__global__ void __launch_bounds__(YIELD_THREADS)
yields_kernel(const YieldArgs a) {
    __shared__ DeviceSpecies sp[YIELD_SPECIES_SMEM];
    extern __shared__ double lam_smem[];        // [ncombo][YIELD_THREADS]
    const int s_begin = blockIdx.y*a.species_per_block;
    const int s_end = min(a.ns, s_begin + a.species_per_block);
    for (int i = threadIdx.x; i < s_end - s_begin; i += blockDim.x) sp[i] = a.species[s_begin + i];
    __syncthreads();

    const int64_t cell = static_cast<int64_t>(blockIdx.x)*blockDim.x + threadIdx.x;
    if (cell >= a.ncell) return;
    const float *__restrict__ S = a.surf;
    const int64_t np = a.ncell_pad;
#define FLD(k) __ldg(&S[static_cast<int64_t>(k)*np + cell])
    const float Tf = FLD(ISS_F_T);
    const float muBf = FLD(ISS_F_MUB), muSf = FLD(ISS_F_MUS), muQf = FLD(ISS_F_MUQ);
    const double temp = Tf;
    const double mu_B = muBf;
    const double dsigma_dot_u = FLD(ISS_F_DA0);
    const double Edec = FLD(ISS_F_E), Pdec = FLD(ISS_F_P), rho_B = FLD(ISS_F_NB);
    const float bulkPif = FLD(ISS_F_BULKPI);

    const ModeFlags &m = a.mode;
    CellCoef cc;
    cell_coefficients(a.tab, m, Edec, rho_B, temp, mu_B, cc);
    if (blockIdx.y == 0) {
        double *co = a.cellcoef + cell*COEF_STRIDE;
#pragma unroll
        for (int i = 0; i < 6; i++) co[i] = cc.c[i];
        co[6] = cc.kappa;
        co[7] = 0.0;
    }

    double bulkPi = 0.0;
    if (m.include_bulk == 1) {
        if (m.kind == 21 || m.kind == 20 || m.kind == 11 || m.kind == 0) {
            bulkPi = bulkPif;                       // GeV/fm^3
        } else {
            bulkPi = static_cast<double>(bulkPif)/HBARC;   // fm^-4
        }
    }

    double dsigma_dot_q = 0.0, prefactor_qmu = 0.0;
    if (m.include_diff == 1) {
        // float arithmetic as in FSSW.cpp:627-629
        const float q = __fadd_rn(__fadd_rn(__fmul_rn(FLD(ISS_F_QX), FLD(ISS_F_DA1)),
                                            __fmul_rn(FLD(ISS_F_QY), FLD(ISS_F_DA2))),
                                  __fmul_rn(FLD(ISS_F_QZ), FLD(ISS_F_DA3)));
        dsigma_dot_q = q;
        prefactor_qmu = rho_B/(Edec + Pdec);
    }
#undef FLD

    const double beta = 1./temp;
    // fugacity of every distinct (B,S,Q): mu in float arithmetic as FSSW.cpp:650
    for (int c = 0; c < a.ncombo; c++) {
        const int4 q = a.combos[c];
        const float muf = __fadd_rn(__fadd_rn(__fmul_rn(static_cast<float>(q.x), muBf),
                                              __fmul_rn(static_cast<float>(q.y), muSf)),
                                    __fmul_rn(static_cast<float>(q.z), muQf));
        lam_smem[c*YIELD_THREADS + threadIdx.x] = exp(beta*static_cast<double>(muf));
    }

    const double unit_factor = 1.0/(HBARC*HBARC*HBARC);
    const bool hot = temp > 0.05;
    const bool bulk_ce = (m.include_bulk == 1) && (m.kind == 1 || m.kind == 21);
    const bool bulk_mom = (m.include_bulk == 1) && (m.kind == 11 || m.kind == 20);
    const bool with_diff = (m.include_diff == 1);
    const SfGrid sf = a.tab.sf;
    const double inv_dx = 1.0/sf.dx;
    const double2 *__restrict__ sf2 = reinterpret_cast<const double2 *>(a.sf4);
    const double common = unit_factor/(2.*M_PI*M_PI);
    // per-cell powers of T: the reference divides by beta = 1/T (FSSW.cpp:812-847); multiplying by
    // T instead differs by one rounding (~1e-16 relative) and removes every division from the
    // species loop
    const double T2 = temp*temp;
    const double T3_over_3 = T2*temp/3.;
    const double sigma = common*dsigma_dot_u;
    const double q_pref = with_diff ? common*dsigma_dot_q/cc.kappa : 0.;

#pragma unroll 1
    for (int is = 0; is < s_end - s_begin; is++) {
        const DeviceSpecies p = sp[is];
        const double mass = p.mass;
        const double lambda = lam_smem[p.combo*YIELD_THREADS + threadIdx.x];
        const int truncate_order = (p.trunc10_mass && hot) ? 10 : 1;
        const double mbeta = mass*beta;
        const double T_over_m = temp*p.inv_mass;

        double N_eq = 0., b1 = 0., b2 = 0., b3 = 0., q2 = 0.;
        double theta = 1.0, fugacity = 1.0;
        for (int n = 1; n <= truncate_order; n++) {
            const double inv_n = (n == 1) ? 1.0 : c_inv_n[n];
            const double arg = n*mass*beta;
            if (n > 1) theta *= -static_cast<double>(p.sign);
            fugacity *= lambda;
            double K_1, K_2, K_3, I_tab;
            if (sf_in_table(sf, arg)) {
                // idx = int((arg - x_min)/dx), frac = remainder/dx (FSSW.cpp:1666-1683); the
                // division by dx is a multiplication here: the lerp is continuous in arg, so a
                // last-bit difference of idx/frac changes the result by ~1e-15 relative
                const double t = (arg - sf.x_min)*inv_dx;
                const int idx = static_cast<int>(t);
                const double frac = t - idx;
                // records idx and idx+1 are adjacent: four 16-byte loads of one 64-byte span
                const double2 k12a = __ldg(sf2 + 2*idx), k3ia = __ldg(sf2 + 2*idx + 1);
                const double2 k12b = __ldg(sf2 + 2*idx + 2), k3ib = __ldg(sf2 + 2*idx + 3);
                const double omf = 1. - frac;
                K_1 = omf*k12a.x + frac*k12b.x;
                K_2 = omf*k12a.y + frac*k12b.y;
                K_3 = omf*k3ia.x + frac*k3ib.x;
                I_tab = omf*k3ia.y + frac*k3ib.y;
            } else {
                bessel_k123(arg, K_1, K_2, K_3);
                I_tab = 0.;
                if (with_diff) {
                    double w[9];
                    expint_weights(w);
#pragma unroll 1
                    for (int i = 0; i < 9; i++) I_tab += w[i]*expint_en(2*i + 2, arg);
                }
            }
            const double tf = theta*fugacity;
            N_eq += tf*inv_n*K_2;
            if (bulk_ce) {
                b1 += tf*(mbeta*K_1 + 3*K_2*inv_n);
                b2 += tf*K_1;
            } else if (bulk_mom) {
                b1 += tf*K_2;
                b2 += tf*(mbeta*K_1 + 3*K_2*inv_n);
                b3 += tf*(mbeta*K_2 + 3*K_3*inv_n);
            }
            if (with_diff) {
                // FSSW.cpp:784-808 with 1/arg = T/(n m)
                const double ra = T_over_m*inv_n;
                const double I_1_n = exp(-arg)*ra*(2.*ra*ra + 2.*ra - 0.5) + I_tab;
                q2 += n*tf*(-(mbeta*mbeta*mbeta)*I_1_n);
            }
        }
        // FSSW.cpp:812-847 and 656-706; the equilibrium series doubles as the first diffusion term
        const double m2 = p.mass2;
        const double pref = p.gspin;
        double total = sigma*m2*temp*N_eq;
        if (m.include_bulk == 1) {
            if (bulk_ce) {
                total += sigma*(-bulkPi*cc.c[0])*(-cc.c[1]*(m2*temp*b1) + m2*mass*(1./3.)*b2);
            } else if (m.kind == 11) {
                total += sigma*bulkPi*((m2*temp*b1)*m2*cc.c[0] + (m2*T2*b2)*p.baryon*cc.c[1]
                                       + (m2*mass*T2*b3)*cc.c[2]);
            } else if (m.kind == 20) {
                total += sigma*bulkPi*((m2*temp*b1)*m2*cc.c[2]
                                       + (m2*T2*b2)*(p.baryon*cc.c[3] + p.strange*cc.c[4]
                                                     + p.charge*cc.c[5])
                                       + (m2*mass*T2*b3)*(cc.c[1] - cc.c[2]));
            }
        }
        if (with_diff) {
            total += q_pref*(-prefactor_qmu*(m2*T2*N_eq) - p.baryon*(T3_over_3*q2));
        }
        a.yields[static_cast<int64_t>(s_begin + is)*np + cell] = fmax(0., pref*total);
    }
}

// K3b: one CTA of 256 threads per (tile of 1024 cells, species); each thread owns 4 consecutive
// cells.  PASS 1 only reduces the tile (tile sums -> K3c); PASS 2 repeats the same fixed-order
// scan, adds the tile base and writes the global inclusive prefix together with level 1 of the
// 16-ary search tree (L1[j] = P[16 j + 15]).  Two reads of the yields instead of read + write +
// read + write of a scan-then-fix-up scheme.
template <int PASS>
__global__ void __launch_bounds__(256)
tile_scan_kernel(const double *__restrict__ yields, double *__restrict__ cdf,
                 double *__restrict__ tilesum, const double *__restrict__ tilebase,
                 double *__restrict__ lev, int64_t lev_stride, int64_t ncell, int64_t ncell_pad,
                 int64_t ntile, int64_t tb_stride, int64_t tb_off) {
    __shared__ double warp_tot[8];
    const int64_t tile = blockIdx.x;
    const int s = blockIdx.y;
    const int64_t base = static_cast<int64_t>(s)*ncell_pad + tile*TILE + threadIdx.x*4;
    const int64_t c0 = tile*TILE + threadIdx.x*4;
    double v[4];
    // ncell_pad is a multiple of TILE and 32-byte aligned rows: vector load
    const double2 a = *reinterpret_cast<const double2 *>(yields + base);
    const double2 b = *reinterpret_cast<const double2 *>(yields + base + 2);
    v[0] = (c0 + 0 < ncell) ? a.x : 0.0;
    v[1] = (c0 + 1 < ncell) ? a.y : 0.0;
    v[2] = (c0 + 2 < ncell) ? b.x : 0.0;
    v[3] = (c0 + 3 < ncell) ? b.y : 0.0;
    v[1] += v[0];
    v[2] += v[1];
    v[3] += v[2];
    double incl = v[3];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const double t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    double offset = incl - v[3];   // exclusive within the warp
    double wbase = 0.0;
    for (int w = 0; w < warp; w++) wbase += warp_tot[w];
    offset += wbase;
    if (PASS == 1) {
        if (threadIdx.x == 255) tilesum[static_cast<int64_t>(s)*ntile + tile] = v[3] + offset;
        return;
    }
    // the same association as the tile-local scan followed by "+ tile base"
    // (surface-chunk mode: tilebase is the GLOBAL table, this handle's tiles start at tb_off)
    const double tb = __ldg(&tilebase[static_cast<int64_t>(s)*tb_stride + tb_off + tile]);
    double2 o0, o1;
    o0.x = (v[0] + offset) + tb; o0.y = (v[1] + offset) + tb;
    o1.x = (v[2] + offset) + tb; o1.y = (v[3] + offset) + tb;
    *reinterpret_cast<double2 *>(cdf + base) = o0;
    *reinterpret_cast<double2 *>(cdf + base + 2) = o1;
    if ((threadIdx.x & 3) == 3)
        lev[static_cast<int64_t>(s)*lev_stride + ((tile*TILE + threadIdx.x*4) >> 4)] = o1.y;
}

// K3c: one warp per species; exclusive prefix over tile sums in a fixed order.
__global__ void tile_base_kernel(const double *__restrict__ tilesum, double *__restrict__ tilebase,
                                 double *__restrict__ total, int64_t ntile) {
    const int s = blockIdx.x;
    const int lane = threadIdx.x;
    double carry = 0.0;
    for (int64_t t0 = 0; t0 < ntile; t0 += 32) {
        const int64_t t = t0 + lane;
        const double v = (t < ntile) ? tilesum[static_cast<int64_t>(s)*ntile + t] : 0.0;
        double incl = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const double x = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += x;
        }
        if (t < ntile) tilebase[static_cast<int64_t>(s)*(ntile + 1) + t] = carry + (incl - v);
        carry += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (lane == 0) {
        tilebase[static_cast<int64_t>(s)*(ntile + 1) + ntile] = carry;
        total[s] = carry;
    }
}

// K3e: level k >= 2 from level k-1 (every 16th entry); entries past the end read as the total
__global__ void cdf_level_kernel(double *__restrict__ lev, const double *__restrict__ total,
                                 int64_t lev_stride, int64_t off_prev, int64_t n_prev,
                                 int64_t off_cur, int64_t n_cur_padded) {
    const int s = blockIdx.y;
    const int64_t j = static_cast<int64_t>(blockIdx.x)*blockDim.x + threadIdx.x;
    if (j >= n_cur_padded) return;
    double *L = lev + static_cast<int64_t>(s)*lev_stride;
    const int64_t src = 16*j + 15;
    L[off_cur + j] = (src < n_prev) ? L[off_prev + src] : __ldg(&total[s]);
}

// pads level 1 beyond its last real entry
__global__ void cdf_level_pad_kernel(double *__restrict__ lev, const double *__restrict__ total,
                                     int64_t lev_stride, int64_t off, int64_t n, int64_t n_padded) {
    const int s = blockIdx.x;
    for (int64_t j = n + threadIdx.x; j < n_padded; j += blockDim.x)
        lev[static_cast<int64_t>(s)*lev_stride + off + j] = __ldg(&total[s]);
}

// K3f: guide table of the cell search (whole-surface mode).  The search needs, for v = (total -
// 1e-15) u, the number of level-1 entries below v (level 1 = the prefix at the end of every 16-cell
// block).  For k = 0..M (M a power of two) G[k] is that number at v_k = (total - 1e-15) (k/M); u in
// [k/M, (k+1)/M) gives v_k <= v <= v_{k+1} (k/M is exact, the product is monotone in its factor), so
// the answer lies in [G[k], G[k+1]]: with M ~ 2 x the number of blocks the bracket holds ~1 entry
// instead of a four-level descent.  One thread per level-1 entry j claims the k with
// L1[j-1] < v_k <= L1[j] (every k in [0, M] has exactly one such j: the last entry claims the rest);
// entries are stored as pairs {G[k], G[k+1]} (one 8-byte load).
__global__ void guide_kernel(const double *__restrict__ lev, int64_t lev_stride, int64_t off1, int64_t n1,
                             const double *__restrict__ total, int2 *__restrict__ guide, int64_t M) {
    const int s = blockIdx.y;
    const int64_t j = static_cast<int64_t>(blockIdx.x)*blockDim.x + threadIdx.x;
    if (j >= n1) return;
    const double *__restrict__ L = lev + static_cast<int64_t>(s)*lev_stride + off1;
    int2 *__restrict__ G = guide + static_cast<int64_t>(s)*(M + 1);
    const double Tp = total[s] - 1e-15;
    const double invM = 1.0/static_cast<double>(M);
    auto claim = [&](int64_t k) {
        G[k].x = static_cast<int>(j);
        if (k > 0) G[k - 1].y = static_cast<int>(j);
    };
    if (!(Tp > 0.)) {
        // empty species: every v is <= 0 <= L1[0], the count is 0 for every k
        for (int64_t k = j*(M + 1)/n1; k < (j + 1)*(M + 1)/n1; k++) G[k] = make_int2(0, 0);
        return;
    }
    const double prev = (j == 0) ? -1.0 : L[j - 1];      // (prefix values are >= 0)
    const double cur = L[j];
    const bool last = (j == n1 - 1);
    if (!(cur > prev) && !last) return;
    int64_t kf = 0;
    if (j > 0) {
        kf = static_cast<int64_t>(floor(prev/Tp*static_cast<double>(M)));
        kf = max(static_cast<int64_t>(0), min(M, kf));
        while (kf > 0 && Tp*(static_cast<double>(kf - 1)*invM) > prev) kf--;
        while (kf <= M && !(Tp*(static_cast<double>(kf)*invM) > prev)) kf++;
    }
    int64_t kl = M;
    if (!last) {
        kl = static_cast<int64_t>(floor(cur/Tp*static_cast<double>(M)));
        kl = max(static_cast<int64_t>(0), min(M, kl));
        while (kl < M && Tp*(static_cast<double>(kl + 1)*invM) <= cur) kl++;
        while (kl >= 0 && !(Tp*(static_cast<double>(kl)*invM) <= cur)) kl--;
    }
    for (int64_t k = kf; k <= kl; k++) claim(k);
    if (last) G[M].y = static_cast<int>(j);
}

// Surface-chunk mode: level 3 of the GLOBAL search tree from the gathered tile sums.  Entry j is the
// prefix value of the last cell of the 4096-cell block j, which is the last value tile 4j+3 writes in
// pass 2: tilesum + tilebase, the same two operands in the same order, hence the same bits as the
// single-GPU level 3 (a copy of that prefix value via levels 1 and 2).
__global__ void chunk_level3_kernel(const double *__restrict__ tilesum, const double *__restrict__ tilebase,
                                    const double *__restrict__ total, int64_t g_ntile, int64_t n1,
                                    double *__restrict__ levg, int64_t levg_stride, int64_t n3_padded) {
    const int s = blockIdx.y;
    const int64_t j = static_cast<int64_t>(blockIdx.x)*blockDim.x + threadIdx.x;
    if (j >= n3_padded) return;
    double v = __ldg(&total[s]);
    if (256*j + 255 < n1) {
        const int64_t t = 4*j + 3;
        v = tilesum[static_cast<int64_t>(s)*g_ntile + t]
            + tilebase[static_cast<int64_t>(s)*(g_ntile + 1) + t];
    }
    levg[static_cast<int64_t>(s)*levg_stride + j] = v;
}

// fills the K_n / E_n grids on the device (FSSW::initialize_special_function_arrays)
__global__ void build_sf_tables_kernel(double *bessel, double *expint, SfGrid g, int with_bulk,
                                       int with_diff) {
    const int i = blockIdx.x*blockDim.x + threadIdx.x;
    if (i >= g.n) return;
    const double x = g.x_min + i*g.dx;
    double k1, k2, k3;
    bessel_k123(x, k1, k2, k3);
    bessel[i*3 + 0] = with_bulk ? k1 : 0.0;
    bessel[i*3 + 1] = k2;
    bessel[i*3 + 2] = with_bulk ? k3 : 0.0;
    if (with_diff) {
        for (int k = 0; k < 9; k++) expint[i*9 + k] = expint_en(2*k + 2, x);
    }
}

// {K1, K2, K3, sum_k w_k E_2k} records the yield kernel reads
__global__ void pack_sf4_kernel(const double *__restrict__ bessel, const double *__restrict__ expint,
                                int n, double *__restrict__ sf4) {
    const int i = blockIdx.x*blockDim.x + threadIdx.x;
    if (i >= n) return;
    double w[9];
    expint_weights(w);
    double acc = 0.;
    if (expint)
        for (int k = 0; k < 9; k++) acc += w[k]*expint[i*9 + k];
    sf4[i*4 + 0] = bessel[i*3 + 0];
    sf4[i*4 + 1] = bessel[i*3 + 1];
    sf4[i*4 + 2] = bessel[i*3 + 2];
    sf4[i*4 + 3] = acc;
}

static int ensure_sf_tables(iss_handle *h, bool need_diff) {
    if (h->d_bessel && (!need_diff || h->d_expint) && h->d_sf4 && h->sf4_with_diff == need_diff)
        return ISS_OK;
    if (!h->d_bessel || (need_diff && !h->d_expint)) {
        // FSSW.cpp:1611-1615
        const double sf_x_min = 0.5, sf_x_max = 400, sf_dx = 0.05;
        SfGrid g;
        g.x_min = sf_x_min;
        g.dx = sf_dx;
        g.x_max_minus_dx = sf_x_max - sf_dx;
        g.n = static_cast<int>((sf_x_max - sf_x_min)/sf_dx) + 1;
        if (!h->d_bessel) ISS_CUDA_TRY(h, cudaMalloc(&h->d_bessel, sizeof(double)*3*g.n));
        if (need_diff && !h->d_expint)
            ISS_CUDA_TRY(h, cudaMalloc(&h->d_expint, sizeof(double)*9*g.n));
        h->sf = g;
        build_sf_tables_kernel<<<(g.n + 127)/128, 128, 0, h->stream>>>(
            h->d_bessel, h->d_expint, g, 1, need_diff ? 1 : 0); ISS_LAUNCHED(h);
        ISS_CUDA_TRY(h, cudaGetLastError());
    }
    if (h->d_sf4) cudaFree(h->d_sf4);
    h->d_sf4 = nullptr;
    ISS_CUDA_TRY(h, cudaMalloc(&h->d_sf4, sizeof(double)*4*(h->sf.n + 1)));
    pack_sf4_kernel<<<(h->sf.n + 127)/128, 128, 0, h->stream>>>(
        h->d_bessel, need_diff ? h->d_expint : nullptr, h->sf.n, h->d_sf4); ISS_LAUNCHED(h);
    ISS_CUDA_TRY(h, cudaGetLastError());
    h->sf4_with_diff = need_diff;
    return ISS_OK;
}

// geometry of the 16-ary search levels over a prefix array of n0 entries (level 0 = the prefix
// itself): level k holds ceil(n_{k-1}/16) entries, stored padded to a multiple of 16
static int level_geometry(int64_t n0, int max_levels, int64_t *lev_n, int64_t *lev_off,
                          int64_t *stride) {
    int64_t n_prev = n0, off = 0;
    int nlev = 0;
    while (n_prev > 16 && nlev < max_levels) {
        const int64_t n = (n_prev + 15)/16;
        const int64_t n_padded = (n + 15)/16*16;
        const int k = ++nlev;
        lev_n[k] = n;
        lev_off[k] = off;
        off += n_padded;
        n_prev = n;
    }
    *stride = off > 0 ? off : 16;
    return nlev;
}

// Part 1 of the yields: per cell x species yields of the cells this handle holds and the sums of
// their 1024-cell tiles.  In surface-chunk mode the caller gathers the tile sums of all ranks
// before part 2.
int run_yields_local(iss_handle *h) {
    if (h->legacy) {
        // the prefix arrays were last sized for the lab-frame surface of the legacy path
        h->ncell = h->ncell_lrf;
        h->ntile = (h->ncell + TILE - 1)/TILE;
        h->ncell_pad = h->ntile*TILE;
        h->legacy = false;
        h->have_yields = false;
    }
    if (h->ncell <= 0 || !h->d_surf) ISS_FAIL(h, ISS_ERR_STATE, "no surface uploaded");
    if (h->nspecies <= 0) ISS_FAIL(h, ISS_ERR_STATE, "no species uploaded");
    if (!h->have_opt) ISS_FAIL(h, ISS_ERR_STATE, "options not set");
    const iss_options &o = h->opt;
    ModeFlags mode;
    mode.include_shear = o.include_deltaf_shear;
    mode.include_bulk = o.include_deltaf_bulk;
    mode.include_diff = o.include_deltaf_diffusion;
    mode.kind = o.bulk_deltaf_kind;
    mode.neos = (o.bulk_deltaf_kind == 21) ? 1 : (o.bulk_deltaf_kind == 20 ? 0 : -1);
    if (mode.neos == 1 && !h->d_ce) ISS_FAIL(h, ISS_ERR_STATE, "CE delta-f table not uploaded");
    if (mode.neos == 0 && !h->d_mom22)
        ISS_FAIL(h, ISS_ERR_STATE, "22-moment delta-f table not uploaded");
    if (mode.include_bulk == 1 && mode.kind == 11 && !h->d_mom14)
        ISS_FAIL(h, ISS_ERR_STATE, "14-moment bulk table not uploaded");
    if (mode.include_diff == 1 && !h->d_kappa)
        ISS_FAIL(h, ISS_ERR_STATE, "kappa_B table not uploaded");
    int rc = ensure_sf_tables(h, o.include_deltaf_diffusion == 1);
    if (rc) return rc;

    h->have_yields = false;
    h->have_local_yields = false;
    h->cellrec_valid = false;
    const int64_t ns = h->nspecies;
    const size_t nval = static_cast<size_t>(ns)*h->ncell_pad;
    ISS_ENSURE(h, h->d_yields, h->yields_bytes, sizeof(double)*nval);
    ISS_ENSURE(h, h->d_cdf, h->cdf_bytes, sizeof(double)*nval);
    ISS_ENSURE(h, h->d_tilesum, h->tilesum_bytes, sizeof(double)*ns*h->ntile);
    ISS_ENSURE(h, h->d_tilebase, h->tilebase_bytes, sizeof(double)*ns*(h->ntile + 1));
    ISS_ENSURE(h, h->d_total, h->total_bytes, sizeof(double)*ns);
    // surface-chunk mode keeps exactly levels 1 and 2 locally (a 4096-cell block = one level-2
    // node; ncell_pad is a multiple of 1024, so both always exist); the levels above them are
    // global (run_yields_finish)
    h->nlev = level_geometry(h->ncell_pad, h->chunk ? 2 : 7, h->lev_n, h->lev_off, &h->lev_stride);
    ISS_ENSURE(h, h->d_cdflev, h->cdflev_bytes, sizeof(double)*ns*h->lev_stride);
    ISS_ENSURE(h, h->d_cellcoef, h->coef_bytes, sizeof(double)*COEF_STRIDE*h->ncell);

    YieldArgs a;
    a.surf = h->d_surf;
    a.ncell = h->ncell;
    a.ncell_pad = h->ncell_pad;
    a.species = h->d_species;
    a.ns = h->nspecies;
    a.tab.bessel = h->d_bessel;
    a.tab.expint = h->d_expint;
    a.tab.sf = h->sf;
    a.tab.ce = h->d_ce;
    a.tab.mom22 = h->d_mom22;
    a.tab.ce_n = h->ce_ne;
    a.tab.mom14 = h->d_mom14;
    a.tab.g14 = h->g14;
    a.tab.kappa = h->d_kappa;
    a.tab.gk = h->gk;
    a.mode = mode;
    a.yields = h->d_yields;
    a.cellcoef = h->d_cellcoef;
    a.sf4 = h->d_sf4;
    a.combos = h->d_combos;
    a.ncombo = h->ncombo;

    const int64_t nblk_x = (h->ncell + YIELD_THREADS - 1)/YIELD_THREADS;
    // enough CTAs to fill 148 SMs several times over, without recomputing the
    // per-cell coefficients more often than needed
    int chunks = 1;
    const int64_t want = 148*16;
    if (nblk_x < want) chunks = static_cast<int>((want + nblk_x - 1)/nblk_x);
    if (chunks > ns) chunks = static_cast<int>(ns);
    int spb = static_cast<int>((ns + chunks - 1)/chunks);
    if (spb > YIELD_SPECIES_SMEM) spb = YIELD_SPECIES_SMEM;
    chunks = static_cast<int>((ns + spb - 1)/spb);
    a.species_per_block = spb;
    {
        ScopedTimer t(h, ISS_T_YIELDS);
        dim3 grid(static_cast<unsigned>(nblk_x), chunks);
        const size_t smem = sizeof(double)*YIELD_THREADS*h->ncombo;
        if (smem + sizeof(DeviceSpecies)*YIELD_SPECIES_SMEM > 48*1024)
            cudaFuncSetAttribute(yields_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 static_cast<int>(smem));
        yields_kernel<<<grid, YIELD_THREADS, smem, h->stream>>>(a); ISS_LAUNCHED(h);
    }
    ISS_CUDA_TRY(h, cudaGetLastError());
    {
        ScopedTimer t(h, ISS_T_SCAN);
        dim3 grid(static_cast<unsigned>(h->ntile), static_cast<unsigned>(ns));
        tile_scan_kernel<1><<<grid, 256, 0, h->stream>>>(h->d_yields, h->d_cdf, h->d_tilesum, nullptr,
                                                         nullptr, 0, h->ncell, h->ncell_pad, h->ntile,
                                                         0, 0); ISS_LAUNCHED(h);
    }
    ISS_CUDA_TRY(h, cudaGetLastError());
    h->have_local_yields = true;
    return ISS_OK;
}

// Part 2: tile bases (fixed-order prefix over ALL tiles of the surface), global inclusive prefix of
// this handle's cells, search levels, species totals.  Surface-chunk mode: the tile sums of the
// whole surface are in h->d_tilesum_g ([ns][g_ntile], assembled by iss_cuda_chunk_yields_finish);
// every rank runs the same tile_base_kernel on the same numbers, so bases, totals and upper search
// levels are bit-identical to a single-GPU run and to each other.
int run_yields_finish(iss_handle *h) {
    if (!h->have_local_yields) ISS_FAIL(h, ISS_ERR_STATE, "yields of the local cells not computed");
    const int64_t ns = h->nspecies;
    {
        ScopedTimer t(h, ISS_T_SCAN);
        dim3 grid(static_cast<unsigned>(h->ntile), static_cast<unsigned>(ns));
        if (h->chunk) {
            ISS_ENSURE(h, h->d_tilebase_g, h->tilebase_g_bytes, sizeof(double)*ns*(h->g_ntile + 1));
            tile_base_kernel<<<static_cast<unsigned>(ns), 32, 0, h->stream>>>(
                h->d_tilesum_g, h->d_tilebase_g, h->d_total, h->g_ntile); ISS_LAUNCHED(h);
            tile_scan_kernel<2><<<grid, 256, 0, h->stream>>>(
                h->d_yields, h->d_cdf, h->d_tilesum, h->d_tilebase_g, h->d_cdflev, h->lev_stride,
                h->ncell, h->ncell_pad, h->ntile, h->g_ntile + 1, h->chunk_tile_begin); ISS_LAUNCHED(h);
        } else {
            tile_base_kernel<<<static_cast<unsigned>(ns), 32, 0, h->stream>>>(
                h->d_tilesum, h->d_tilebase, h->d_total, h->ntile); ISS_LAUNCHED(h);
            tile_scan_kernel<2><<<grid, 256, 0, h->stream>>>(
                h->d_yields, h->d_cdf, h->d_tilesum, h->d_tilebase, h->d_cdflev, h->lev_stride,
                h->ncell, h->ncell_pad, h->ntile, h->ntile + 1, 0); ISS_LAUNCHED(h);
        }
        if (h->nlev >= 1) {
            const int64_t n1p = (h->lev_n[1] + 15)/16*16;
            if (n1p > h->lev_n[1]) {
                cdf_level_pad_kernel<<<static_cast<unsigned>(ns), 32, 0, h->stream>>>(
                    h->d_cdflev, h->d_total, h->lev_stride, h->lev_off[1], h->lev_n[1], n1p); ISS_LAUNCHED(h);
            }
        }
        for (int k = 2; k <= h->nlev; k++) {
            const int64_t np_ = (h->lev_n[k] + 15)/16*16;
            dim3 g3(static_cast<unsigned>((np_ + 127)/128), static_cast<unsigned>(ns));
            cdf_level_kernel<<<g3, 128, 0, h->stream>>>(h->d_cdflev, h->d_total, h->lev_stride,
                                                        h->lev_off[k - 1], h->lev_n[k - 1],
                                                        h->lev_off[k], np_); ISS_LAUNCHED(h);
        }
        h->guide_M = 0;
        if (!h->chunk && h->nlev >= 1) {
            // guide table of the cell search: M = power of two >= 2 x (number of 16-cell blocks)
            int64_t M = 16;
            while (M < 2*h->lev_n[1] && M < (int64_t(1) << 26)) M <<= 1;
            ISS_ENSURE(h, h->d_guide, h->guide_bytes, sizeof(int2)*static_cast<size_t>(ns)*(M + 1));
            dim3 gg(static_cast<unsigned>((h->lev_n[1] + 127)/128), static_cast<unsigned>(ns));
            guide_kernel<<<gg, 128, 0, h->stream>>>(h->d_cdflev, h->lev_stride, h->lev_off[1], h->lev_n[1],
                                                    h->d_total, static_cast<int2 *>(h->d_guide), M);
            ISS_LAUNCHED(h);
            h->guide_M = M;
        }
        if (h->chunk) {
            // global levels 3..g_nlev in their own array (offsets relative to level 3)
            int64_t gn[8] = {0}, goff[8] = {0}, gstride = 0;
            h->g_nlev = level_geometry(h->g_ntile*TILE, 7, gn, goff, &gstride);
            if (h->g_nlev < 3)
                ISS_FAIL(h, ISS_ERR_ARG, "surface-chunk mode needs a surface of more than 4096 cells");
            for (int k = 3; k <= h->g_nlev; k++) {
                h->g_lev_n[k] = gn[k];
                h->g_lev_off[k] = goff[k] - goff[3];
            }
            h->g_lev_stride = gstride - goff[3];
            ISS_ENSURE(h, h->d_cdflev_g, h->cdflev_g_bytes, sizeof(double)*ns*h->g_lev_stride);
            const int64_t n3p = (gn[3] + 15)/16*16;
            dim3 g3(static_cast<unsigned>((n3p + 127)/128), static_cast<unsigned>(ns));
            chunk_level3_kernel<<<g3, 128, 0, h->stream>>>(h->d_tilesum_g, h->d_tilebase_g, h->d_total,
                                                           h->g_ntile, gn[1], h->d_cdflev_g,
                                                           h->g_lev_stride, n3p); ISS_LAUNCHED(h);
            for (int k = 4; k <= h->g_nlev; k++) {
                const int64_t np_ = (gn[k] + 15)/16*16;
                dim3 g4(static_cast<unsigned>((np_ + 127)/128), static_cast<unsigned>(ns));
                cdf_level_kernel<<<g4, 128, 0, h->stream>>>(h->d_cdflev_g, h->d_total, h->g_lev_stride,
                                                            h->g_lev_off[k - 1], gn[k - 1],
                                                            h->g_lev_off[k], np_); ISS_LAUNCHED(h);
            }
        }
    }
    ISS_CUDA_TRY(h, cudaGetLastError());
    h->h_total.resize(ns);
    ISS_CUDA_TRY(h, cudaMemcpyAsync(h->h_total.data(), h->d_total, sizeof(double)*ns,
                                    cudaMemcpyDeviceToHost, h->stream));
    ISS_CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    h->have_yields = true;
    h->lambda_on_device = false;
    return ISS_OK;
}

int legacy_args(iss_handle *h, LegacyArgs &G) {
    if (!h->have_legopt) ISS_FAIL(h, ISS_ERR_STATE, "iss_cuda_legacy_set_options must run first");
    if (h->nlab <= 0 || !h->d_lab) ISS_FAIL(h, ISS_ERR_STATE, "no lab-frame surface uploaded");
    if (h->nlegpos != h->nlab || !h->d_legpos)
        ISS_FAIL(h, ISS_ERR_STATE, "cell positions of the lab-frame surface not uploaded");
    if (!h->d_zx) ISS_FAIL(h, ISS_ERR_STATE, "z_exp_m_z table not uploaded");
    const iss_legacy_options &o = h->legopt;
    if (o.include_deltaf_bulk == 1 && o.bulk_deltaf_kind == 0 && !h->d_bulk0)
        ISS_FAIL(h, ISS_ERR_STATE, "legacy sampler: bulk_deltaf_kind 0 needs the 14-moment coefficient "
                                   "table (iss_cuda_upload_table, ISS_TABLE_BULK14)");
    if (o.include_deltaf_diffusion == 1 && !h->d_kappa)
        ISS_FAIL(h, ISS_ERR_STATE, "kappa_B table not uploaded");
    if (!(o.sample_pT_up_to > 0.) || !(o.sample_y_minus_eta_s_range > 0.))
        ISS_FAIL(h, ISS_ERR_ARG, "legacy sampler: pT and y - eta_s ranges must be positive");
    if (!h->d_lambert) {
        ISS_CUDA_TRY(h, cudaMalloc(&h->d_lambert, sizeof(double)*LEGACY_LAMBERT_N));
        legacy_lambert_kernel<<<(LEGACY_LAMBERT_N + 255)/256, 256, 0, h->stream>>>(h->d_lambert);
        ISS_LAUNCHED(h);
        ISS_CUDA_TRY(h, cudaGetLastError());
    }
    G.lab = h->d_lab;
    G.pos = h->d_legpos;
    G.coef = h->d_legcoef;
    G.bulk0 = h->d_bulk0;
    G.nbulk0 = h->nbulk0;
    G.zx = h->d_zx;
    G.zy = h->d_zy;
    G.nz = h->nz;
    G.lambert = h->d_lambert;
    G.include_shear = o.include_deltaf_shear;
    G.include_bulk = o.include_deltaf_bulk;
    G.bulk_kind = o.bulk_deltaf_kind;
    G.include_diff = o.include_deltaf_diffusion;
    G.restrict_deltaf = o.restrict_deltaf;
    G.deltaf_max_ratio = o.deltaf_max_ratio;
    G.pT_to = o.sample_pT_up_to;
    G.y_range = o.sample_y_minus_eta_s_range;
    G.ncell = h->nlab;
    G.ncell_pad = h->ncell_pad;
    G.tab.bessel = h->d_bessel;
    G.tab.expint = h->d_expint;
    G.tab.sf = h->sf;
    G.tab.ce = nullptr;
    G.tab.mom22 = nullptr;
    G.tab.ce_n = 0;
    G.tab.mom14 = nullptr;
    G.tab.g14 = h->g14;
    G.tab.kappa = h->d_kappa;
    G.tab.gk = h->gk;
    G.yields = h->d_yields;
    G.max_out = nullptr;
    return ISS_OK;
}

// Yields of the legacy path over the lab-frame surface (h->d_lab), then the same fixed-order
// prefix / search levels / totals as the FSSW path (run_yields_finish).
int run_legacy_yields(iss_handle *h, double *yields_host, double *maximum_host) {
    if (h->nspecies <= 0) ISS_FAIL(h, ISS_ERR_STATE, "no species uploaded");
    if (!h->have_opt) ISS_FAIL(h, ISS_ERR_STATE, "options not set");
    if (!h->have_legopt) ISS_FAIL(h, ISS_ERR_STATE, "iss_cuda_legacy_set_options must run first");
    // K_n / E_n tables: built on the device unless the caller uploaded them
    int rc = ensure_sf_tables(h, h->legopt.include_deltaf_diffusion == 1);
    if (rc) return rc;
    const int64_t ncell = h->nlab;
    if (ncell <= 0) ISS_FAIL(h, ISS_ERR_STATE, "no lab-frame surface uploaded");
    if (ncell >= (int64_t(1) << 31)) ISS_FAIL(h, ISS_ERR_ARG, "ncell must be < 2^31");
    // the FSSW surface (if any) is superseded: the prefix arrays now describe the lab-frame cells
    h->ncell = ncell;
    h->ntile = (ncell + TILE - 1)/TILE;
    h->ncell_pad = h->ntile*TILE;
    h->chunk = false;
    h->have_yields = false;
    h->have_local_yields = false;
    h->cellrec_valid = false;
    h->have_batch = false;
    h->legacy = false;
    const int64_t ns = h->nspecies;
    const size_t nval = static_cast<size_t>(ns)*h->ncell_pad;
    ISS_ENSURE(h, h->d_yields, h->yields_bytes, sizeof(double)*nval);
    ISS_ENSURE(h, h->d_cdf, h->cdf_bytes, sizeof(double)*nval);
    ISS_ENSURE(h, h->d_tilesum, h->tilesum_bytes, sizeof(double)*ns*h->ntile);
    ISS_ENSURE(h, h->d_tilebase, h->tilebase_bytes, sizeof(double)*ns*(h->ntile + 1));
    ISS_ENSURE(h, h->d_total, h->total_bytes, sizeof(double)*ns);
    h->nlev = level_geometry(h->ncell_pad, 7, h->lev_n, h->lev_off, &h->lev_stride);
    ISS_ENSURE(h, h->d_cdflev, h->cdflev_bytes, sizeof(double)*ns*h->lev_stride);
    ISS_ENSURE(h, h->d_legcoef, h->legcoef_bytes, sizeof(double4)*ncell);
    LegacyArgs G;
    rc = legacy_args(h, G);
    if (rc) return rc;
    ISS_CUDA_TRY(h, cudaMemsetAsync(h->d_yields, 0, sizeof(double)*nval, h->stream));
    {
        ScopedTimer t(h, ISS_T_YIELDS);
        legacy_coef_kernel<<<static_cast<unsigned>((ncell + 127)/128), 128, 0, h->stream>>>(G);
        ISS_LAUNCHED(h);
        legacy_yields_kernel<<<static_cast<unsigned>((ncell + 127)/128), 128, 0, h->stream>>>(
            G, h->d_species, static_cast<int>(ns)); ISS_LAUNCHED(h);
    }
    ISS_CUDA_TRY(h, cudaGetLastError());
    if (yields_host)
        ISS_CUDA_TRY(h, cudaMemcpy2DAsync(yields_host, sizeof(double)*ncell, h->d_yields,
                                          sizeof(double)*h->ncell_pad, sizeof(double)*ncell, ns,
                                          cudaMemcpyDeviceToHost, h->stream));
    if (maximum_host) {
        ISS_ENSURE(h, h->d_legmax, h->legmax_bytes, sizeof(double)*ns*ncell);
        G.max_out = h->d_legmax;
        const int64_t n = ns*ncell;
        legacy_max_kernel<<<static_cast<unsigned>((n + 127)/128), 128, 0, h->stream>>>(
            G, h->d_species, static_cast<int>(ns)); ISS_LAUNCHED(h);
        ISS_CUDA_TRY(h, cudaGetLastError());
        ISS_CUDA_TRY(h, cudaMemcpyAsync(maximum_host, h->d_legmax, sizeof(double)*n,
                                        cudaMemcpyDeviceToHost, h->stream));
    }
    {
        ScopedTimer t(h, ISS_T_SCAN);
        legacy_clamp_kernel<<<static_cast<unsigned>((nval + 255)/256), 256, 0, h->stream>>>(
            h->d_yields, static_cast<int64_t>(nval)); ISS_LAUNCHED(h);
        dim3 grid(static_cast<unsigned>(h->ntile), static_cast<unsigned>(ns));
        tile_scan_kernel<1><<<grid, 256, 0, h->stream>>>(h->d_yields, h->d_cdf, h->d_tilesum, nullptr,
                                                         nullptr, 0, h->ncell, h->ncell_pad, h->ntile,
                                                         0, 0); ISS_LAUNCHED(h);
    }
    ISS_CUDA_TRY(h, cudaGetLastError());
    h->have_local_yields = true;
    rc = run_yields_finish(h);
    if (rc) return rc;
    h->legacy = true;
    return ISS_OK;
}

int run_yields(iss_handle *h) {
    if (h->chunk)
        ISS_FAIL(h, ISS_ERR_STATE, "surface-chunk mode: use iss_cuda_chunk_yields_local/_finish");
    int rc = run_yields_local(h);
    if (rc) return rc;
    return run_yields_finish(h);
}

}  // namespace iss
