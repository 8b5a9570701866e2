// gpu_fssw.cpp -- see gpu_fssw.h.  Citations are to the reference's src/FSSW.cpp.
#include "gpu_fssw.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <map>
#include <mutex>
#include <sstream>
#include <thread>
#include <unordered_map>

#include "engine_pool.h"
#include "logger.h"
#include "writers.h"

using iss_host::info;

namespace {

double seconds_since(const std::chrono::steady_clock::time_point &t0) {
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

// ISS_PROFILE=1: wall-clock of the host-side phases on stderr
struct PhaseTimer {
    const char *name;
    std::chrono::steady_clock::time_point t0;
    static bool on() {
        static int v = -1;
        if (v < 0) {
            const char *e = getenv("ISS_PROFILE");
            v = (e && atoi(e) == 1) ? 1 : 0;
        }
        return v == 1;
    }
    explicit PhaseTimer(const char *n) : name(n), t0(std::chrono::steady_clock::now()) {}
    ~PhaseTimer() {
        if (on()) fprintf(stderr, "[iss profile] %-28s %8.3f ms\n", name, 1e3*seconds_since(t0));
    }
};

// numbers of a whitespace separated text file after `skip_lines` header lines
std::vector<double> read_numbers(const std::string &file, int skip_lines) {
    std::vector<double> out;
    FILE *f = fopen(file.c_str(), "r");
    if (!f) {
        iss_host::error("Can not found file: " + file);
        exit(1);
    }
    fseek(f, 0, SEEK_END);
    const long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    std::vector<char> buf(n + 1);
    const size_t got = fread(buf.data(), 1, n, f);
    fclose(f);
    buf[got] = '\0';
    char *p = buf.data();
    for (int i = 0; i < skip_lines && p; i++) {
        p = strchr(p, '\n');
        if (p) p++;
    }
    if (!p) return out;
    out.reserve(got/12);
    for (;;) {
        char *e;
        const double v = strtod(p, &e);
        if (e == p) break;
        out.push_back(v);
        p = e;
    }
    return out;
}

// Process-wide caches.  MUSIC/JETSCAPE-style hosts call iSS::generate_samples() once per
// hydro event and the reference builds a fresh FSSW every time (iSS.cpp:145-149); rebuilding
// the CUDA context objects, re-pinning gigabytes of host memory and re-parsing the coefficient
// tables on every call would dominate the run, so they are kept for the life of the process.
struct HandlePool {
    std::mutex mu;
    std::map<int, std::vector<iss_handle *>> idle;     // by device
};
HandlePool &handle_pool() {
    // never destroyed: CUDA runtime calls during static destruction (after main() has returned)
    // are undefined, so the pooled handles are left to the driver at process exit, like the
    // pinned blocks below
    static HandlePool *p = new HandlePool();
    return *p;
}

using iss_pool::PinnedBlock;
struct PinnedPool {
    std::mutex mu;
    std::vector<PinnedBlock> idle;
};
PinnedPool &pinned_pool() {
    static PinnedPool p;      // blocks are left to the driver at process exit
    return p;
}

// a pinned block of at least `bytes` (the largest idle one if it fits, else a new allocation)
PinnedBlock pool_acquire(iss_handle *h, int64_t bytes) {
    PinnedPool &P = pinned_pool();
    {
        std::lock_guard<std::mutex> lk(P.mu);
        for (size_t i = 0; i < P.idle.size(); i++)
            if (P.idle[i].bytes >= bytes) {
                PinnedBlock b = P.idle[i];
                P.idle.erase(P.idle.begin() + i);
                return b;
            }
    }
    PinnedBlock b;
    if (iss_cuda_host_alloc(h, &b.ptr, bytes) != ISS_OK) {
        iss_host::error(std::string("iss_cuda_host_alloc failed: ") + iss_cuda_last_error(h));
        exit(-1);
    }
    b.bytes = bytes;
    return b;
}

void pool_release(iss_handle *h, PinnedBlock b) {
    if (!b.ptr) return;
    PinnedPool &P = pinned_pool();
    std::lock_guard<std::mutex> lk(P.mu);
    // keep at most two idle blocks (hadron list + surface staging): drop the smallest
    P.idle.push_back(b);
    while (P.idle.size() > 2) {
        size_t k = 0;
        for (size_t i = 1; i < P.idle.size(); i++)
            if (P.idle[i].bytes < P.idle[k].bytes) k = i;
        iss_cuda_host_free(h, P.idle[k].ptr);
        P.idle.erase(P.idle.begin() + k);
    }
}

const std::vector<double> &cached_numbers_impl(const std::string &file, int skip_lines) {
    static std::mutex mu;
    static std::map<std::string, std::vector<double>> cache;
    std::lock_guard<std::mutex> lk(mu);
    const std::string key = file + "#" + std::to_string(skip_lines);
    auto it = cache.find(key);
    if (it == cache.end()) it = cache.emplace(key, read_numbers(file, skip_lines)).first;
    return it->second;
}

const std::vector<double> &cached_numbers(const std::string &file, int skip_lines) {
    return cached_numbers_impl(file, skip_lines);
}

}  // namespace

namespace iss_pool {

int default_device() {
    int device = 0;
    if (const char *lr = getenv("LOCAL_RANK")) device = atoi(lr);
    if (const char *dv = getenv("ISS_CUDA_DEVICE")) device = atoi(dv);
    return device;
}

iss_handle *acquire_handle(int device) {
    iss_handle *h = nullptr;
    {
        HandlePool &P = handle_pool();
        std::lock_guard<std::mutex> lk(P.mu);
        auto &v = P.idle[device];
        if (!v.empty()) {
            h = v.back();
            v.pop_back();
        }
    }
    if (!h && iss_cuda_create(device, &h) != ISS_OK) {
        iss_host::error("iss_cuda_create: no usable CUDA device (the B200 engine has no CPU "
                        "fallback)");
        exit(-1);
    }
    return h;
}

void release_handle(int device, iss_handle *h) {
    if (!h) return;
    HandlePool &P = handle_pool();
    std::lock_guard<std::mutex> lk(P.mu);
    P.idle[device].push_back(h);
}

const std::vector<double> &cached_numbers(const std::string &file, int skip_lines) {
    return cached_numbers_impl(file, skip_lines);
}

PinnedBlock pinned_acquire(iss_handle *h, int64_t bytes) { return pool_acquire(h, bytes); }
void pinned_release(iss_handle *h, PinnedBlock b) { pool_release(h, b); }

}  // namespace iss_pool

void GpuFSSW::check_(int rc, const char *what) {
    if (rc == ISS_OK) return;
    std::ostringstream os;
    os << what << " failed (status " << rc << "): " << (h_ ? iss_cuda_last_error(h_) : "no handle");
    iss_host::error(os.str());
    exit(rc == ISS_ERR_RANGE ? 1 : -1);
}

GpuFSSW::GpuFSSW(long seed, const std::vector<int> &chosen_monvals,
                 const std::vector<particle_info> &particles,
                 const std::vector<FO_surf_LRF> &FOsurf_LRF, int flag_PCE,
                 ParameterReader *paraRdr, std::string path, std::string table_path,
                 AfterburnerType afterburner_type, const float *packed_lrf)
    : paraRdr_(paraRdr), path_(path), table_path_(table_path),
      afterburner_type_(afterburner_type), particles_(particles), surf_(FOsurf_LRF),
      packed_lrf_(packed_lrf), seed_(seed) {
    init_(chosen_monvals, flag_PCE);
}

namespace {
const std::vector<FO_surf_LRF> k_no_lrf_surface;
}

GpuFSSW::GpuFSSW(long seed, const std::vector<int> &chosen_monvals,
                 const std::vector<particle_info> &particles,
                 const std::vector<FO_surf> &FOsurf_lab, int flag_PCE, ParameterReader *paraRdr,
                 std::string path, std::string table_path, AfterburnerType afterburner_type)
    : paraRdr_(paraRdr), path_(path), table_path_(table_path),
      afterburner_type_(afterburner_type), particles_(particles), surf_(k_no_lrf_surface),
      lab_surf_(&FOsurf_lab), legacy_(true), seed_(seed) {
    init_(chosen_monvals, flag_PCE);
}

void GpuFSSW::init_(const std::vector<int> &chosen_monvals, int flag_PCE) {
    if (flag_PCE != 0) {
        iss_host::error("partial chemical equilibrium EoS is not supported by the B200 engine");
        exit(1);
    }
    // keys consumed on the FSSW path (FSSW.cpp:67-103)
    hydro_mode_ = static_cast<int>(paraRdr_->getVal("hydro_mode"));
    use_oscar_ = static_cast<int>(paraRdr_->getVal("use_OSCAR_format"));
    use_gzip_ = static_cast<int>(paraRdr_->getVal("use_gzip_format"));
    use_binary_ = static_cast<int>(paraRdr_->getVal("use_binary_format"));
    include_shear_ = static_cast<int>(paraRdr_->getVal("include_deltaf_shear"));
    include_bulk_ = static_cast<int>(paraRdr_->getVal("include_deltaf_bulk"));
    bulk_kind_ = static_cast<int>(paraRdr_->getVal("bulk_deltaf_kind"));
    include_diff_ = static_cast<int>(paraRdr_->getVal("include_deltaf_diffusion"));
    number_of_repeated_sampling_ =
        static_cast<int>(paraRdr_->getVal("number_of_repeated_sampling"));
    (void)paraRdr_->getVal("maximum_sampling_events");
    (void)paraRdr_->getVal("output_samples_into_files");
    flag_perform_decays_ = (paraRdr_->getVal("perform_decays") == 1) ? 1 : 0;
    const int spectator_mode = static_cast<int>(paraRdr_->getVal("include_spectators", 0));
    // spectators are an FSSW feature (FSSW.cpp:349-350); the legacy class has none
    flag_spectators_ = (spectator_mode != 0 && !legacy_) ? 1 : 0;
    if (flag_spectators_) read_spectators_(path_ + "/spectators.dat");
    if (legacy_) {
        // per-species samples_<monval>.dat / samples_control_<monval>.dat / samples_format.dat
        // (emissionfunction.cpp:3353-3375, 3441-3560, 3578-3617); the samples are kept in memory
        // as well, whatever store_samples_in_memory says
        flag_sample_files_ = (paraRdr_->getVal("output_samples_into_files") == 1) ? 1 : 0;
        if (paraRdr_->getVal("store_samples_in_memory") != 1)
            iss_host::warning("MC_sampling = 2: samples are always kept in memory by the B200 engine");
    }

    device_ = iss_pool::default_device();
    h_ = iss_pool::acquire_handle(device_);
    { PhaseTimer t("select_species"); select_species_(chosen_monvals); }
    { PhaseTimer t("upload_surface"); if (legacy_) upload_lab_surface_(); else upload_surface_(); }
    { PhaseTimer t("upload_tables"); upload_tables_(); }
    { PhaseTimer t("upload_decay_table"); upload_decay_table_(); }

    iss_options opt;
    memset(&opt, 0, sizeof(opt));
    opt.hydro_mode = hydro_mode_;
    opt.include_deltaf_shear = include_shear_;
    opt.include_deltaf_bulk = include_bulk_;
    opt.include_deltaf_diffusion = include_diff_;
    opt.bulk_deltaf_kind = bulk_kind_;
    opt.dN_dy_sampling_model = static_cast<int>(paraRdr_->getVal("dN_dy_sampling_model"));
    opt.dN_dy_sampling_para1 = paraRdr_->getVal("dN_dy_sampling_para1");
    opt.local_charge_conservation =
        static_cast<int>(paraRdr_->getVal("local_charge_conservation"));
    opt.y_LB = paraRdr_->getVal("y_LB");
    opt.y_RB = paraRdr_->getVal("y_RB");
    if (opt.dN_dy_sampling_model > 100) {
        std::cout << "FSSW::sample_using_dN_dxtdy_4all_particles error: sampling model "
                  << opt.dN_dy_sampling_model << " is not supported." << std::endl;
        exit(-1);
    }
    check_(iss_cuda_set_options(h_, &opt), "iss_cuda_set_options");
    if (legacy_) {
        iss_legacy_options lo;
        memset(&lo, 0, sizeof(lo));
        lo.include_deltaf_shear = include_shear_;
        lo.include_deltaf_bulk = include_bulk_;
        lo.bulk_deltaf_kind = bulk_kind_;
        lo.include_deltaf_diffusion = include_diff_;
        lo.restrict_deltaf = static_cast<int>(paraRdr_->getVal("restrict_deltaf"));
        lo.deltaf_max_ratio = paraRdr_->getVal("deltaf_max_ratio");
        // emissionfunction.cpp:3302-3306
        double pT_to = paraRdr_->getVal("sample_pT_up_to");
        if (pT_to < 0) {
            const std::string pT_file = table_path_ + "/bin_tables/pT_gauss_table.dat";
            const std::vector<double> &pT_tab = cached_numbers(pT_file, 0);
            if (pT_tab.size() < 2) {
                iss_host::error("Can not found file: " + pT_file);
                exit(1);
            }
            pT_to = pT_tab[pT_tab.size() - 2];      // first column of the last row
        }
        lo.sample_pT_up_to = pT_to;
        lo.sample_y_minus_eta_s_range = paraRdr_->getVal("sample_y_minus_eta_s_range");
        // TableFunction z_exp_m_z (emissionfunction.cpp:3284-3287)
        const std::string zfile = table_path_ + "/z_exp_m_z.dat";
        const std::vector<double> &z = cached_numbers(zfile, 0);
        if (z.size() < 8) {
            iss_host::error("Can not found file: " + zfile);
            exit(1);
        }
        std::vector<double> zx(z.size()/2), zy(z.size()/2);
        for (size_t i = 0; i < zx.size(); i++) {
            zx[i] = z[2*i];
            zy[i] = z[2*i + 1];
        }
        check_(iss_cuda_legacy_upload_z_table(h_, zx.data(), zy.data(), static_cast<int>(zx.size())),
               "iss_cuda_legacy_upload_z_table");
        check_(iss_cuda_legacy_set_options(h_, &lo), "iss_cuda_legacy_set_options");
    }
}

// FO_surf -> the ISS_L_* record of include/iss_cuda.h and the cell positions (legacy mode)
void GpuFSSW::upload_lab_surface_() {
    const std::vector<FO_surf> &surf = *lab_surf_;
    const int64_t n = static_cast<int64_t>(surf.size());
    std::vector<float> rec(static_cast<size_t>(n)*ISS_LAB_NFIELD), pos(static_cast<size_t>(n)*4);
    for (int64_t l = 0; l < n; l++) {
        const FO_surf &c = surf[l];
        float *r = rec.data() + l*ISS_LAB_NFIELD;
        r[ISS_L_TAU] = c.tau;
        r[ISS_L_U0] = c.u0; r[ISS_L_U1] = c.u1; r[ISS_L_U2] = c.u2; r[ISS_L_U3] = c.u3;
        r[ISS_L_DA0] = c.da0; r[ISS_L_DA1] = c.da1; r[ISS_L_DA2] = c.da2; r[ISS_L_DA3] = c.da3;
        r[ISS_L_T] = c.Tdec; r[ISS_L_P] = c.Pdec; r[ISS_L_E] = c.Edec;
        r[ISS_L_MUB] = c.muB; r[ISS_L_MUS] = c.muS; r[ISS_L_MUQ] = c.muQ;
        r[ISS_L_PI00] = c.pi00; r[ISS_L_PI01] = c.pi01; r[ISS_L_PI02] = c.pi02; r[ISS_L_PI03] = c.pi03;
        r[ISS_L_PI11] = c.pi11; r[ISS_L_PI12] = c.pi12; r[ISS_L_PI13] = c.pi13;
        r[ISS_L_PI22] = c.pi22; r[ISS_L_PI23] = c.pi23; r[ISS_L_PI33] = c.pi33;
        r[ISS_L_BULKPI] = c.bulkPi; r[ISS_L_BN] = c.Bn;
        r[ISS_L_Q0] = c.qmu0; r[ISS_L_Q1] = c.qmu1; r[ISS_L_Q2] = c.qmu2; r[ISS_L_Q3] = c.qmu3;
        r[ISS_L_SPARE] = 0.f;
        float *q = pos.data() + l*4;
        q[0] = c.xpt; q[1] = c.ypt; q[2] = c.eta; q[3] = 0.f;
    }
    check_(iss_cuda_upload_surface_lab(h_, rec.data(), n), "iss_cuda_upload_surface_lab");
    check_(iss_cuda_legacy_upload_positions(h_, pos.data(), n), "iss_cuda_legacy_upload_positions");
}

GpuFSSW::~GpuFSSW() {
    if (!h_) return;
    iss_cuda_fetch_wait(h_);
    iss_cuda_synchronize(h_);
    PinnedBlock b;
    b.ptr = hadrons_;
    b.bytes = hadron_cap_*static_cast<int64_t>(sizeof(iSS_Hadron));
    pool_release(h_, b);
    iss_pool::release_handle(device_, h_);
}

// chosen list -> indices into the pdg table, unknown ids dropped with a warning, then a stable
// ascending sort by mass (the reference's bubble sort only swaps on strict >, FSSW.cpp:115-162)
std::vector<int> GpuFSSW::order_species(const std::vector<int> &chosen_monvals,
                                        const std::vector<particle_info> &particles,
                                        bool sort_by_mass) {
    std::vector<int> order, missing;
    for (int monval : chosen_monvals) {
        int found = -1;
        for (size_t n = 0; n < particles.size(); n++)
            if (particles[n].monval == monval) {
                found = static_cast<int>(n);
                break;
            }
        if (found >= 0) order.push_back(found);
        else missing.push_back(monval);
    }
    if (!missing.empty()) {
        iss_host::warning("not all chosen particles are in the pdg particle list!");
        std::ostringstream os;
        os << "There are " << missing.size() << " particles can not be found in the pdg particle list!";
        iss_host::warning(os.str());
        iss_host::warning("Their monte carlo numbers are:");
        for (int m : missing) iss_host::warning(std::to_string(m));
    }
    if (sort_by_mass)
        std::stable_sort(order.begin(), order.end(),
                         [&](int a, int b) { return particles[a].mass < particles[b].mass; });
    return order;
}

void GpuFSSW::select_species_(const std::vector<int> &chosen_monvals) {
    // the legacy class sorts only when grouping_particles is set (emissionfunction.cpp:190-206)
    const bool sort_by_mass = !legacy_ || paraRdr_->getVal("grouping_particles") != 0;
    species_table_idx_ = order_species(chosen_monvals, particles_, sort_by_mass);
    for (int idx : species_table_idx_) {
        const particle_info &p = particles_[idx];
        iss_species s;
        memset(&s, 0, sizeof(s));
        s.pid = p.monval;
        s.gspin = p.gspin;
        s.baryon = p.baryon;
        s.strange = p.strange;
        s.charge = p.charge;
        s.sign = p.sign;
        s.decay_idx = idx;      // the decay table is uploaded in pdg-table order
        s.mass = p.mass;
        species_.push_back(s);
    }
    check_(iss_cuda_upload_species(h_, species_.data(), static_cast<int>(species_.size())),
           "iss_cuda_upload_species");
}

void GpuFSSW::pack_surface(const std::vector<FO_surf_LRF> &surf, float *dst, int64_t c0, int64_t c1) {
    auto pack = [&](int64_t b, int64_t e) {
        for (int64_t c = b; c < e; c++) {
            const FO_surf_LRF &s = surf[c];
            float *r = dst + c*ISS_NFIELD;
            r[0] = s.tau; r[1] = s.xpt; r[2] = s.ypt; r[3] = s.eta;
            for (int k = 0; k < 4; k++) {
                r[4 + k] = s.da_mu_LRF[k];
                r[8 + k] = s.u_tz[k];
            }
            r[12] = s.Edec; r[13] = s.Tdec; r[14] = s.Pdec; r[15] = s.Bn;
            r[16] = s.muB; r[17] = s.muS; r[18] = s.muQ; r[19] = s.bulkPi;
            r[20] = s.piLRF_xx; r[21] = s.piLRF_xy; r[22] = s.piLRF_xz; r[23] = s.piLRF_yy;
            r[24] = s.piLRF_yz; r[25] = s.qmuLRF_x; r[26] = s.qmuLRF_y; r[27] = s.qmuLRF_z;
        }
    };
    const int64_t n = c1 - c0;
    const int nthread = static_cast<int>(std::max<int64_t>(
        1, std::min<int64_t>(16, std::min<int64_t>(std::thread::hardware_concurrency(), n/65536))));
    if (nthread <= 1) {
        pack(c0, c1);
        return;
    }
    std::vector<std::thread> pool;
    for (int t = 0; t < nthread; t++) pool.emplace_back(pack, c0 + n*t/nthread, c0 + n*(t + 1)/nthread);
    for (auto &t : pool) t.join();
}

void GpuFSSW::upload_surface_() {
    const int64_t n = static_cast<int64_t>(surf_.size());
    if (packed_lrf_) {
        // the caller holds the records in upload layout in pinned memory: one copy, no staging
        check_(iss_cuda_upload_surface_aos(h_, packed_lrf_, n), "iss_cuda_upload_surface_aos");
        return;
    }
    PinnedBlock stage = pool_acquire(h_, n*ISS_NFIELD*static_cast<int64_t>(sizeof(float)));
    float *dst = static_cast<float *>(stage.ptr);
    // in parts: while one part travels to the device (and is transposed there) the next is packed
    const int nparts = (n >= 262144) ? 4 : 1;
    for (int part = 0; part < nparts; part++) {
        const int64_t p0 = n*part/nparts, p1 = n*(part + 1)/nparts;
        pack_surface(surf_, dst, p0, p1);
        check_(iss_cuda_upload_surface_aos_part(h_, dst + p0*ISS_NFIELD, p0, p1 - p0, n),
               "iss_cuda_upload_surface_aos_part");
    }
    pool_release(h_, stage);
}

// delta-f coefficient tables, file formats of FSSW.cpp:1215-1376 and 1546-1568
void GpuFSSW::upload_tables_() {
    const bool smash = (afterburner_type_ == AfterburnerType::SMASH);
    const std::string dir = table_path_ + "/deltaf_tables";
    if (!legacy_ && include_bulk_ == 1 && bulk_kind_ == 11) {
        // c0.dat, c1.dat, c2.dat: "101", "81", one text header, then rows "T muB value" with T
        // running fastest.  All three files have THREE header lines; the reference skips only two
        // for c0 and ends up with an unfilled c0 table -- parsed correctly here.
        const std::string folder = dir + (smash ? "/smash_box" : "/urqmd");
        std::vector<double> tab;
        double grid[4] = {0, 0, 0, 0};
        int nT = 0, nmu = 0;
        for (int c = 0; c < 3; c++) {
            const std::string file = folder + "/c" + std::to_string(c) + ".dat";
            const std::vector<double> &head = cached_numbers(file, 0);
            if (head.size() < 2) {
                iss_host::error("Can not found file: " + file);
                exit(1);
            }
            nT = static_cast<int>(head[0]);
            nmu = static_cast<int>(head[1]);
            const std::vector<double> &v = cached_numbers(file, 3);
            if (static_cast<long>(v.size()) < 3L*nT*nmu) {
                iss_host::error("short 14-moment table: " + file);
                exit(1);
            }
            if (c == 0) {
                tab.assign(static_cast<size_t>(3)*nT*nmu, 0.);
                grid[0] = v[0];                 // T0
                grid[1] = v[3] - v[0];          // dT   (second row, FSSW.cpp:1285-1288)
                grid[2] = v[1];                 // mu0
                grid[3] = v[3*nT + 1] - v[1];   // dmu  (first row of the second mu block)
            }
            for (int j = 0; j < nmu; j++)
                for (int i = 0; i < nT; i++)
                    tab[(static_cast<size_t>(c)*nT + i)*nmu + j] = v[3*(static_cast<size_t>(j)*nT + i) + 2];
        }
        check_(iss_cuda_upload_table(h_, ISS_TABLE_MOM14, tab.data(), nT, nmu, grid),
               "iss_cuda_upload_table(14-moment)");
    }
    if (legacy_) {
        // polynomial bulk coefficients for kinds 1-4 (emissionfunction.cpp:3625-3762); kind 0 reads
        // "T[1/fm] B0 D0 E0" rows (emissionfunction.cpp:298-301)
        if (include_bulk_ == 1 && bulk_kind_ == 0) {
            const std::string file = dir + "/BulkDf_Coefficients_Hadrons_s95p-v0-PCE.dat";
            const std::vector<double> &v = cached_numbers(file, 0);
            if (v.size() < 16 || v.size() % 4 != 0) {
                iss_host::error("Can not found file: " + file);
                exit(1);
            }
            check_(iss_cuda_upload_table(h_, ISS_TABLE_BULK14, v.data(), static_cast<int64_t>(v.size()/4),
                                         4, nullptr),
                   "iss_cuda_upload_table(14-moment bulk, kind 0)");
        }
    } else if (bulk_kind_ == 21) {
        const std::string file = dir + (smash ? "/smash" : "/urqmd") + "/NEoSBQS_CE_deltafCoeff.dat";
        const std::vector<double> &v = cached_numbers(file, 1);
        if (v.size() < 200u*200u*5u) {
            iss_host::error("short CE table: " + file);
            exit(1);
        }
        check_(iss_cuda_upload_table(h_, ISS_TABLE_CE, v.data(), 200, 200, nullptr),
               "iss_cuda_upload_table(CE)");
    } else if (bulk_kind_ == 20) {
        const std::string file =
            dir + (smash ? "/smash" : "/urqmd") + "/NEoSBQS_22mom_deltafCoeff.dat";
        const std::vector<double> &v = cached_numbers(file, 1);
        if (v.size() < 200u*200u*8u) {
            iss_host::error("short 22-moment table: " + file);
            exit(1);
        }
        check_(iss_cuda_upload_table(h_, ISS_TABLE_MOM22, v.data(), 200, 200, nullptr),
               "iss_cuda_upload_table(22-moment)");
    }
    if (include_diff_ == 1) {
        // 100 (mu_B) x 150 (T) rows "T muB kappa", T fastest; the grid is hard-coded in the
        // reference (FSSW.cpp:1548-1553)
        const std::string file = dir + "/Coefficients_RTA_diffusion.dat";
        const std::vector<double> &v = cached_numbers(file, 0);
        const int nT = 150, nmu = 100;
        if (static_cast<long>(v.size()) < 3L*nT*nmu) {
            iss_host::error("short kappa_B table: " + file);
            exit(1);
        }
        std::vector<double> tab(static_cast<size_t>(nT)*nmu);
        for (int j = 0; j < nmu; j++)
            for (int i = 0; i < nT; i++)
                tab[static_cast<size_t>(i)*nmu + j] = v[3*(static_cast<size_t>(j)*nT + i) + 2];
        const double grid[4] = {0.05, 0.001, 0.0, 0.007892};
        check_(iss_cuda_upload_table(h_, ISS_TABLE_KAPPA_B, tab.data(), nT, nmu, grid),
               "iss_cuda_upload_table(kappa_B)");
    }
}

// the decay table is the whole pdg list (particle_decay.cpp:33-172 re-reads the same file);
// it is uploaded always because the QA kernel takes quantum numbers from it.
void GpuFSSW::upload_decay_table_() {
    std::vector<iss_decay_species> sp(particles_.size());
    std::vector<iss_decay_channel> ch;
    std::unordered_map<int, int> row_of;    // first row with a given Monte-Carlo id
    for (size_t i = 0; i < particles_.size(); i++) row_of.emplace(particles_[i].monval, static_cast<int>(i));
    for (size_t i = 0; i < particles_.size(); i++) {
        const particle_info &p = particles_[i];
        iss_decay_species &d = sp[i];
        memset(&d, 0, sizeof(d));
        d.pid = p.monval;
        d.stable = p.stable;
        d.n_channels = p.decays;
        d.first_channel = static_cast<int>(ch.size());
        d.baryon = p.baryon;
        d.strange = p.strange;
        d.charge = p.charge;
        d.mass = p.mass;
        d.width = p.width;
        for (int j = 0; j < p.decays; j++) {
            iss_decay_channel c;
            memset(&c, 0, sizeof(c));
            c.n_part = p.decay_channels[j]->decay_Npart;
            c.branching_ratio = p.decay_channels[j]->branching_ratio;
            for (int k = 0; k < 5; k++) {
                c.daughter[k] = -1;
                const int monval = p.decay_channels[j]->decay_part[k];
                if (monval == 0) continue;
                const auto it = row_of.find(monval);
                if (it != row_of.end()) c.daughter[k] = it->second;
            }
            ch.push_back(c);
        }
    }
    check_(iss_cuda_upload_decay_table(h_, sp.data(), static_cast<int>(sp.size()), ch.data(),
                                       static_cast<int>(ch.size())),
           "iss_cuda_upload_decay_table");
}

// spectators.dat: one comment line, rows "t x y z mass px py rapidity charge" (Spectators.cpp:17-52)
void GpuFSSW::read_spectators_(const std::string &file) {
    std::ifstream in(file.c_str());
    if (!in.good()) {
        std::cout << "[Error]: Spectator file : " << file << " not found!" << std::endl;
        exit(1);
    }
    std::string line;
    std::getline(in, line);
    while (std::getline(in, line)) {
        if (in.eof()) break;    // an unterminated last line is not used by the reference
        std::stringstream ss(line);
        double t, x, y, z, mass, px, py, rap;
        int e_charge = 0;
        if (!(ss >> t >> x >> y >> z >> mass >> px >> py >> rap >> e_charge)) continue;
        iSS_Hadron hd;
        hd.pid = (e_charge == 0) ? 2112 : 2212;
        hd.mass = static_cast<float>(mass);
        hd.t = static_cast<float>(t);
        hd.x = static_cast<float>(x);
        hd.y = static_cast<float>(y);
        hd.z = static_cast<float>(z);
        const double mT = std::sqrt(mass*mass + px*px + py*py);
        hd.E = static_cast<float>(mT*std::cosh(rap));
        hd.px = static_cast<float>(px);
        hd.py = static_cast<float>(py);
        hd.pz = static_cast<float>(mT*std::sinh(rap));
        spectators_.push_back(hd);
    }
}

void GpuFSSW::reserve_hadrons_(int64_t need) {
    if (need <= hadron_cap_) return;
    int64_t cap = std::max<int64_t>(need, hadron_cap_ + hadron_cap_/2);
    cap = std::max<int64_t>(cap, 1024);
    PinnedBlock b = pool_acquire(h_, cap*static_cast<int64_t>(sizeof(iSS_Hadron)));
    if (hadrons_) {
        check_(iss_cuda_fetch_wait(h_), "iss_cuda_fetch_wait");     // copies into the old block
        memcpy(b.ptr, hadrons_, sizeof(iSS_Hadron)*event_off_.back());
        PinnedBlock old;
        old.ptr = hadrons_;
        old.bytes = hadron_cap_*static_cast<int64_t>(sizeof(iSS_Hadron));
        pool_release(h_, old);
    }
    hadrons_ = static_cast<iSS_Hadron *>(b.ptr);
    hadron_cap_ = b.bytes/static_cast<int64_t>(sizeof(iSS_Hadron));
}

void GpuFSSW::compute_yields() {
    PhaseTimer t("compute_yields");
    dN_species_.assign(species_.size(), 0.);
    if (legacy_)
        check_(iss_cuda_legacy_compute_yields(h_, dN_species_.data(), nullptr, nullptr),
               "iss_cuda_legacy_compute_yields");
    else
        check_(iss_cuda_compute_yields(h_, dN_species_.data(), nullptr), "iss_cuda_compute_yields");
}

// FSSW::compute_number_of_sampling_needed (FSSW.cpp:851-869): uses pdg-table entry 1 as "pi+"
int GpuFSSW::compute_number_of_sampling_needed_(long number_of_particles_needed) {
    double dNdy_thermal_pion = -1.;
    for (size_t n = 0; n < species_table_idx_.size(); n++)
        if (species_table_idx_[n] == 1) dNdy_thermal_pion = dN_species_[n];
    if (dNdy_thermal_pion < 0. && particles_.size() > 1 && !legacy_) {
        // the reference evaluates pdg-table entry 1 whatever the chosen list holds
        // (FSSW.cpp:851-869): one extra yield pass over that species alone, then the chosen
        // list is put back
        const particle_info &p = particles_[1];
        iss_species one;
        memset(&one, 0, sizeof(one));
        one.pid = p.monval; one.gspin = p.gspin; one.baryon = p.baryon; one.strange = p.strange;
        one.charge = p.charge; one.sign = p.sign; one.decay_idx = 1; one.mass = p.mass;
        check_(iss_cuda_upload_species(h_, &one, 1), "iss_cuda_upload_species");
        double dN1 = 0.;
        check_(iss_cuda_compute_yields(h_, &dN1, nullptr), "iss_cuda_compute_yields");
        dNdy_thermal_pion = dN1;
        check_(iss_cuda_upload_species(h_, species_.data(), static_cast<int>(species_.size())),
               "iss_cuda_upload_species");
        compute_yields();
    }
    if (!(dNdy_thermal_pion > 0.)) {
        iss_host::error("sample_upto_desired_particle_number: the thermal yield of pdg-table entry 1 "
                        "is not positive, can not derive the number of events");
        exit(-1);
    }
    // int(...) as in the reference, evaluated in 64 bits so that large requests do not overflow
    long nev = static_cast<long>(std::min(1.0e15, number_of_particles_needed/(6.*dNdy_thermal_pion)));
    if (hydro_mode_ == 2) nev *= 10;
    // the legacy class caps at 10000 events (emissionfunction.cpp:3268)
    const long max_ev = legacy_ ? 10000 : static_cast<long>(paraRdr_->getVal("maximum_sampling_events"));
    return static_cast<int>(std::max<long>(1, std::min(max_ev, nev)));
}

void GpuFSSW::sample_events() {
    const auto t0 = std::chrono::steady_clock::now();
    info(" Function sample_using_dN_dxtdy_4all_particles started...");
    if (dN_species_.empty()) compute_yields();
    if (static_cast<int>(paraRdr_->getVal("sample_upto_desired_particle_number")) == 1) {
        number_of_repeated_sampling_ = compute_number_of_sampling_needed_(
            static_cast<long>(paraRdr_->getVal("number_of_particles_needed")));
    }
    info("Sampling using dN/dy with sample_using_dN_dxtdy_4all_particles function.");
    info("number of repeated sampling = " + std::to_string(number_of_repeated_sampling_));
    const double y_LB = paraRdr_->getVal("y_LB"), y_RB = paraRdr_->getVal("y_RB");
    double dN_event = 0.;
    // the reference prints two lines per species here (FSSW.cpp:960-962); 642 lines per call are
    // noise in front of a 50 ms call, so they appear only with ISS_VERBOSE=1 (or AMOUNT_OF_OUTPUT > 0)
    static const bool per_species_log = (iSS_data::AMOUNT_OF_OUTPUT > 0)
                                        || (getenv("ISS_VERBOSE") && atoi(getenv("ISS_VERBOSE")) > 0);
    for (size_t n = 0; n < species_.size(); n++) {
        const particle_info &p = particles_[species_table_idx_[n]];
        const double dN_dy = dN_species_[n];
        const double dN = (hydro_mode_ != 2) ? (y_RB - y_LB)*dN_dy : dN_dy;
        dN_event += dN;
        if (!per_species_log) continue;
        std::ostringstream os;
        os << "Index: " << n << ", Name: " << p.name << ", Monte-carlo index: " << p.monval;
        info(os.str());
        std::ostringstream os2;
        os2 << " -- Sampling using dN_dy=" << dN_dy << ", dN=" << dN << "...";
        info(os2.str());
    }
    {
        std::ostringstream os;
        os << " -- Sampling " << species_.size() << " species, dN per event = " << dN_event;
        info(os.str());
    }

    nev_ = number_of_repeated_sampling_;
    event_off_.assign(1, 0);
    event_cache_.clear();
    event_cache_.resize(nev_);

    // events per batch: bounded by the free device memory (40 B per hadron in two output buffers,
    // x4 head-room for decays) and small enough that the device->host copy of one batch overlaps
    // the sampling of the next (at least 16 batches for large runs)
    int64_t free_b = 0, total_b = 0;
    check_(iss_cuda_mem_info(h_, &free_b, &total_b), "iss_cuda_mem_info");
    // device bytes per primary hadron: two 40-B output buffers + 48-B task (+ decays: counts,
    // decayed list ~1.6x longer) with head-room for Poisson fluctuations
    const double per_event = std::max(1.0, dN_event)*(flag_perform_decays_ ? 320.0 : 150.0)
                             + 24.0*species_.size();
    int64_t batch = static_cast<int64_t>(0.5*static_cast<double>(free_b)/per_event);
    const double hadrons_total = dN_event*static_cast<double>(nev_);
    int64_t min_batches = 16;       // measured on the C4 step: 8 -> 50.0 ms, 16 -> 49.0 ms, 32 -> 52.9 ms per call
    if (const char *e = getenv("ISS_BATCHES")) min_batches = std::max(1, atoi(e));
    if (hadrons_total > 4e6) batch = std::min<int64_t>(batch, (nev_ + min_batches - 1)/min_batches);
    batch = std::max<int64_t>(1, std::min<int64_t>(batch, nev_));
    // FSSW::shell skips the feed-down for SMASH (FSSW.cpp:346); EmissionFunctionArray::shell does
    // not (emissionfunction.cpp:2570-2572)
    const bool decays_on = flag_perform_decays_ && (legacy_ || afterburner_type_ != AfterburnerType::SMASH);
    if (decays_on) std::cout << "perform resonance decays... " << std::endl;
    const int32_t qa_pids[2] = {211, 2212};
    const int64_t nsp = static_cast<int64_t>(spectators_.size());
    reserve_hadrons_(static_cast<int64_t>(dN_event*nev_*(decays_on ? 2.0 : 1.02)
                                         + 6.0*std::sqrt(dN_event*nev_ + 1.0))
                     + nsp*nev_ + 1024);

    // engine addition: `first_event_index` (default 0) numbers this call's events
    // first, first + 1, ...  Every random stream is keyed by (seed, event index, ...), so processes
    // that share the seed and use disjoint index ranges (one per GPU) produce together exactly the
    // events a single process would produce for the whole range.
    const int64_t ev_base = static_cast<int64_t>(paraRdr_->getValQuiet("first_event_index", 0));
    if (ev_base < 0) {
        iss_host::error("first_event_index must be >= 0");
        exit(-1);
    }
    PhaseTimer tb("batches (sample+copy)");
    if (flag_sample_files_) {
        begin_sample_files_();
        check_(iss_cuda_set_trace(h_, 1), "iss_cuda_set_trace");
    }
    for (int64_t ev0 = 0; ev0 < nev_; ev0 += batch) {
        const int64_t ev1 = std::min<int64_t>(nev_, ev0 + batch);
        iss_counts cnt;
        check_(iss_cuda_sample(h_, static_cast<uint64_t>(seed_), ev_base + ev0, ev_base + ev1, &cnt),
               "iss_cuda_sample");
        if (flag_sample_files_) append_sample_files_(ev1 - ev0, cnt.n_hadrons);
        if (decays_on) {
            // FSSW::shell skips the feed-down for SMASH (FSSW.cpp:346)
            check_(iss_cuda_decay(h_, static_cast<uint64_t>(seed_), &cnt), "iss_cuda_decay");
        }
        check_(iss_cuda_histograms(h_, qa_pids, 2, ev0 > 0 ? 1 : 0), "iss_cuda_histograms");
        std::vector<int64_t> off(ev1 - ev0 + 1);
        check_(iss_cuda_event_offsets(h_, off.data()), "iss_cuda_event_offsets");
        const int64_t base = event_off_.back();
        if (nsp == 0) {
            reserve_hadrons_(base + cnt.n_hadrons);
            int64_t got = 0;
            check_(iss_cuda_fetch_all_async(h_, reinterpret_cast<iss_hadron *>(hadrons_ + base),
                                            hadron_cap_ - base, &got),
                   "iss_cuda_fetch_all_async");
            for (int64_t i = 1; i <= ev1 - ev0; i++) event_off_.push_back(base + off[i]);
        } else {
            // spectators are appended to every event (FSSW::addSpectatorsToHadronList)
            reserve_hadrons_(base + cnt.n_hadrons + nsp*(ev1 - ev0));
            int64_t pos = base;
            for (int64_t i = 0; i < ev1 - ev0; i++) {
                int64_t got = 0;
                check_(iss_cuda_fetch_event(h_, i, reinterpret_cast<iss_hadron *>(hadrons_ + pos),
                                            hadron_cap_ - pos, &got),
                       "iss_cuda_fetch_event");
                pos += got;
                memcpy(hadrons_ + pos, spectators_.data(), sizeof(iSS_Hadron)*nsp);
                pos += nsp;
                event_off_.push_back(pos);
            }
        }
    }
    { PhaseTimer tw("final fetch_wait"); check_(iss_cuda_fetch_wait(h_), "iss_cuda_fetch_wait"); }
    if (flag_sample_files_) {
        end_sample_files_();
        check_(iss_cuda_set_trace(h_, 0), "iss_cuda_set_trace");
    }
    if (flag_spectators_) std::cout << "Add spectators to the hadron list... " << std::endl;
    if (static_cast<int>(paraRdr_->getValQuiet("reduce_checks_over_ranks", 0)) == 1) {
        // one process per GPU, events sharded over the ranks (first_event_index): the QA block
        // behind iSS::perform_checks is summed over the ranks; hadron lists stay rank-local
        join_ranks_();
        check_(iss_cuda_histograms_allreduce(h_, nullptr), "iss_cuda_histograms_allreduce");
    }
    qa_.assign(iss_cuda_qa_size(), 0.);
    check_(iss_cuda_qa_fetch(h_, qa_.data()), "iss_cuda_qa_fetch");
    if (nsp > 0) add_spectators_to_qa_(qa_pids, 2);
    std::cout << std::endl
              << "sample_using_dN_dxtdy_4all_particles finished in " << seconds_since(t0)
              << " seconds." << std::endl;
}

// Rank and size of the job: parameters nccl_rank / nccl_nranks, else the environment of the
// launcher (torchrun: RANK / WORLD_SIZE, Open MPI, PMI, Slurm).  Rendezvous through a file: rank 0
// writes the 128-byte NCCL id to $ISS_NCCL_ID_FILE (default /tmp/iss_nccl_id_<MASTER_PORT>), the
// others wait for it; ncclCommInitRank returns once every rank has joined, then rank 0 removes
// the file.  The communicator lives on the pooled handle for the rest of the process.
void GpuFSSW::join_ranks_() {
    auto env_int = [](std::initializer_list<const char *> names, int fallback) {
        for (const char *n : names)
            if (const char *v = getenv(n)) return atoi(v);
        return fallback;
    };
    int nranks = static_cast<int>(paraRdr_->getValQuiet("nccl_nranks", -1));
    int rank = static_cast<int>(paraRdr_->getValQuiet("nccl_rank", -1));
    if (nranks < 0)
        nranks = env_int({"WORLD_SIZE", "OMPI_COMM_WORLD_SIZE", "PMI_SIZE", "SLURM_NTASKS"}, 1);
    if (rank < 0) rank = env_int({"RANK", "OMPI_COMM_WORLD_RANK", "PMI_RANK", "SLURM_PROCID"}, 0);
    qa_ranks_ = std::max(1, nranks);
    if (nranks <= 1) return;
    static std::mutex mu;
    static std::map<iss_handle *, int> joined;      // handle -> size of the communicator it holds
    std::lock_guard<std::mutex> lk(mu);
    if (joined.count(h_) && joined[h_] == nranks) return;
    std::string file;
    if (const char *f = getenv("ISS_NCCL_ID_FILE")) file = f;
    else file = std::string("/tmp/iss_nccl_id_") + (getenv("MASTER_PORT") ? getenv("MASTER_PORT") : "0");
    unsigned char id[128];
    if (rank == 0) {
        if (iss_cuda_nccl_unique_id(id) != ISS_OK) {
            iss_host::error("reduce_checks_over_ranks: NCCL (libnccl.so.2) is not available");
            exit(-1);
        }
        const std::string tmp = file + ".tmp";
        FILE *f = fopen(tmp.c_str(), "wb");
        if (!f || fwrite(id, 1, sizeof(id), f) != sizeof(id)) {
            iss_host::error("reduce_checks_over_ranks: can not write " + tmp);
            exit(-1);
        }
        fclose(f);
        rename(tmp.c_str(), file.c_str());
    } else {
        const auto t0 = std::chrono::steady_clock::now();
        for (;;) {
            FILE *f = fopen(file.c_str(), "rb");
            if (f) {
                const size_t got = fread(id, 1, sizeof(id), f);
                fclose(f);
                if (got == sizeof(id)) break;
            }
            if (seconds_since(t0) > 300.) {
                iss_host::error("reduce_checks_over_ranks: no NCCL id in " + file + " after 300 s");
                exit(-1);
            }
            std::this_thread::sleep_for(std::chrono::milliseconds(20));
        }
    }
    check_(iss_cuda_nccl_init(h_, id, rank, nranks), "iss_cuda_nccl_init");
    if (rank == 0) remove(file.c_str());
    joined[h_] = nranks;
}

// ---- per-species sample files of the legacy class (output_samples_into_files = 1) ---------------
// One text line per PRIMARY hadron in samples_<monval>.dat, events one after the other, and the
// number of draws of every event in samples_control_<monval>.dat.  Quirks kept: under local charge
// conservation the conjugate partner is written into the positive species' file while the control
// file counts the draws (emissionfunction.cpp:3441-3546); negative species get empty files only
// if they were reached before the `continue` (they are not: no files for them, :3343-3349).
void GpuFSSW::begin_sample_files_() {
    sample_files_.assign(species_.size(), nullptr);
    control_files_.assign(species_.size(), nullptr);
    const int lcc = static_cast<int>(paraRdr_->getVal("local_charge_conservation"));
    for (size_t s = 0; s < species_.size(); s++) {
        const std::string a = path_ + "/samples_control_" + std::to_string(species_[s].pid) + ".dat";
        const std::string b = path_ + "/samples_" + std::to_string(species_[s].pid) + ".dat";
        if (lcc == 1 && species_[s].charge < 0) continue;       // skipped before the files are made
        remove(a.c_str());
        remove(b.c_str());
        control_files_[s] = fopen(a.c_str(), "w");
        sample_files_[s] = fopen(b.c_str(), "w");
        if (!control_files_[s] || !sample_files_[s]) {
            iss_host::error("can not open " + b);
            exit(-1);
        }
    }
}

void GpuFSSW::append_sample_files_(int64_t nev_batch, int64_t n_hadrons) {
    const int ns = static_cast<int>(species_.size());
    const int lcc = static_cast<int>(paraRdr_->getVal("local_charge_conservation"));
    std::vector<int64_t> mult(static_cast<size_t>(nev_batch)*ns);
    check_(iss_cuda_get_multiplicities(h_, mult.data()), "iss_cuda_get_multiplicities");
    std::vector<iss_hadron> had(static_cast<size_t>(std::max<int64_t>(n_hadrons, 1)));
    std::vector<int32_t> cell(had.size()), tries(had.size());
    int64_t got = 0;
    check_(iss_cuda_fetch_all(h_, had.data(), static_cast<int64_t>(had.size()), &got), "iss_cuda_fetch_all");
    if (got > 0) check_(iss_cuda_get_trace(h_, cell.data(), tries.data()), "iss_cuda_get_trace");
    const std::vector<FO_surf> &surf = *lab_surf_;
    int64_t pos = 0;
    for (int64_t ev = 0; ev < nev_batch; ev++)
        for (int s = 0; s < ns; s++) {
            const int64_t n = mult[ev*ns + s];
            const int64_t nout = (lcc == 1 && species_[s].charge > 0) ? 2*n : n;
            if (!control_files_[s]) continue;       // (negative species under charge pairing: n = 0)
            fprintf(control_files_[s], "%lu\n", static_cast<unsigned long>(n));
            for (int64_t i = 0; i < nout; i++, pos++) {
                const iss_hadron &h = had[pos];
                const FO_surf &c = surf[cell[pos]];
                const double px = h.px, py = h.py, pz = h.pz, E = h.E, t = h.t, z = h.z;
                if (use_oscar_) {
                    fprintf(sample_files_[s],
                            "%24.16e  %24.16e  %24.16e  %24.16e  %24.16e  %24.16e  %24.16e  %24.16e  %24.16e\n",
                            px, py, pz, E, static_cast<double>(h.mass), static_cast<double>(h.x),
                            static_cast<double>(h.y), z, t);
                    continue;
                }
                const double pT = std::sqrt(px*px + py*py);
                double phi = std::atan2(py, px);
                if (phi < 0.) phi += 2.*M_PI;           // the reference samples phi in [0, 2 pi)
                const double rap = 0.5*std::log((E + pz)/(E - pz));
                const double eta_s = 0.5*std::log((t + z)/(t - z));
                fprintf(sample_files_[s],
                        "%lu  %e  %e  %e  %e  %e  %e  %e  %e  %e  %e  %e  %e  %e  %e  %e  %e  %e\n",
                        static_cast<unsigned long>(cell[pos]), c.tau, c.xpt, c.ypt, rap - eta_s, pT, phi,
                        c.da0, c.da1, c.da2, c.u1/c.u0, c.u2/c.u0, rap, eta_s, E, pz, t, z);
            }
        }
}

void GpuFSSW::end_sample_files_() {
    for (FILE *f : sample_files_) if (f) fclose(f);
    for (FILE *f : control_files_) if (f) fclose(f);
    sample_files_.clear();
    control_files_.clear();
    // emissionfunction.cpp:3578-3617 (ParameterReader lower-cases the keys when it reads them back)
    std::ofstream of((path_ + "/samples_format.dat").c_str());
    if (!use_oscar_) {
        of << "Total_number_of_columns = " << 18 << std::endl << "FZ_cell_idx = " << 1 << std::endl
           << "tau = " << 2 << std::endl << "FZ_x = " << 3 << std::endl << "FZ_y = " << 4 << std::endl
           << "y_minus_eta_s = " << 5 << std::endl << "pT = " << 6 << std::endl << "phi = " << 7 << std::endl
           << "surf_da0 = " << 8 << std::endl << "surf_da1 = " << 9 << std::endl << "surf_da2 = " << 10
           << std::endl << "surf_vx = " << 11 << std::endl << "surf_vy = " << 12 << std::endl
           << "y = " << 13 << std::endl << "eta_s = " << 14 << std::endl << "E = " << 15 << std::endl
           << "p_z = " << 16 << std::endl << "t = " << 17 << std::endl << "z = " << 18 << std::endl;
    } else {
        of << "Total_number_of_columns = " << 9 << std::endl << "t = " << 9 << std::endl
           << "FZ_x = " << 6 << std::endl << "FZ_y = " << 7 << std::endl << "z = " << 8 << std::endl
           << "E = " << 4 << std::endl << "px = " << 1 << std::endl << "py = " << 2 << std::endl
           << "p_z = " << 3 << std::endl << "mass =" << 5 << std::endl;
    }
}

void GpuFSSW::shell() {
    PhaseTimer t("shell total");
    compute_yields();
    { PhaseTimer t2("sample_events"); sample_events(); }
    if (!legacy_) computeAvgTotalEnergyMomentum();
    // priority OSCAR > gzip > binary (FSSW.cpp:354-360)
    if (use_oscar_) combine_samples_to_OSCAR();
    else if (use_gzip_) combine_samples_to_gzip_file();
    else if (use_binary_) combine_samples_to_binary_file();
}

std::vector<iSS_Hadron> *GpuFSSW::get_hadron_list_iev(const int iev) {
    if (iev < 0 || iev >= nev_) {
        iss_host::error("get_hadron_list_iev: event index out of range");
        exit(-1);
    }
    if (!event_cache_[iev]) {
        event_cache_[iev].reset(new std::vector<iSS_Hadron>(hadrons_ + event_off_[iev],
                                                            hadrons_ + event_off_[iev + 1]));
    }
    return event_cache_[iev].get();
}

// The QA block is accumulated on the device from the sampled hadrons; the reference runs its checks
// over Hadron_list AFTER addSpectatorsToHadronList (FSSW.cpp:349-352, iSS.cpp:59-83), so the same
// spectators, appended to every event, are folded in here: with a per-event constant c added to a
// per-event quantity a,  sum (a + c) = S1 + n c  and  sum (a + c)^2 = S2 + 2 c S1 + n c^2.
// Binning as in qa_kernel (iss_b200/csrc/qa.cu).  n = events behind the block (all ranks' when it
// was reduced; every rank appends the same spectators).
void GpuFSSW::add_spectators_to_qa_(const int32_t *pids, int npid) {
    const double n = qa_[0];
    double sp[4] = {0, 0, 0, 0}, T[16] = {0}, net[3] = {0, 0, 0};
    for (const iSS_Hadron &hd : spectators_) {
        const double p[4] = {hd.E, hd.px, hd.py, hd.pz};
        for (int a = 0; a < 4; a++) {
            sp[a] += p[a];
            for (int c = 0; c < 4; c++) T[4*a + c] += p[a]*p[c]/p[0];
        }
        net[0] += 1.;                               // nucleons
        net[2] += (hd.pid == 2212) ? 1. : 0.;
    }
    for (int i = 0; i < 4; i++) {
        const double S1 = qa_[1 + i];
        qa_[1 + i] = S1 + n*sp[i];
        qa_[5 + i] = qa_[5 + i] + 2.*sp[i]*S1 + n*sp[i]*sp[i];
    }
    for (int i = 0; i < 16; i++) qa_[9 + i] += n*T[i];
    qa_[25] += n*static_cast<double>(spectators_.size());
    for (int i = 0; i < 3; i++) qa_[26 + i] += n*net[i];
    for (int k = 0; k < npid; k++) {
        double *blk = qa_.data() + ISS_QA_HEAD + static_cast<size_t>(k)*ISS_QA_PER;
        double c_pt[ISS_QA_NPT] = {0}, s_pt[ISS_QA_NPT] = {0};
        double c_tot = 0.;
        for (const iSS_Hadron &hd : spectators_) {
            if (hd.pid != pids[k]) continue;
            c_tot += 1.;
            const double pT = std::sqrt(static_cast<double>(hd.px)*hd.px + static_cast<double>(hd.py)*hd.py);
            const int ib = static_cast<int>(pT/(5.0/(ISS_QA_NPT - 1)));
            if (ib >= 0 && ib < ISS_QA_NPT) {
                c_pt[ib] += 1.;
                s_pt[ib] += pT;
            }
            const double y = std::asinh(hd.pz/std::sqrt(static_cast<double>(hd.mass)*hd.mass + pT*pT));
            const int iy = static_cast<int>(std::floor((y + 5.0)/(10.0/ISS_QA_NY)));
            if (iy >= 0 && iy < ISS_QA_NY) blk[3*ISS_QA_NPT + iy] += n;
            const double phi = std::atan2(static_cast<double>(hd.py), static_cast<double>(hd.px));
            int iphi = static_cast<int>(std::floor((phi + M_PI)/(2.*M_PI/ISS_QA_NPHI)));
            iphi = std::min(ISS_QA_NPHI - 1, std::max(0, iphi));
            blk[3*ISS_QA_NPT + ISS_QA_NY + iphi] += n;
            const int iv = static_cast<int>(pT/(3.0/ISS_QA_NV2));
            if (iv < ISS_QA_NV2) {
                const double c2 = (pT > 0.) ? (static_cast<double>(hd.px)*hd.px
                                               - static_cast<double>(hd.py)*hd.py)/(pT*pT) : 0.;
                blk[3*ISS_QA_NPT + ISS_QA_NY + ISS_QA_NPHI + iv] += n*c2;
                blk[3*ISS_QA_NPT + ISS_QA_NY + ISS_QA_NPHI + ISS_QA_NV2 + iv] += n;
            }
        }
        for (int ib = 0; ib < ISS_QA_NPT; ib++) {
            if (c_pt[ib] == 0.) continue;
            const double cnt = blk[ib];
            blk[2*ISS_QA_NPT + ib] += 2.*c_pt[ib]*cnt + n*c_pt[ib]*c_pt[ib];
            blk[ib] = cnt + n*c_pt[ib];
            blk[ISS_QA_NPT + ib] += n*s_pt[ib];
        }
        if (c_tot > 0.) {
            const double tot = blk[ISS_QA_PER - 2];
            blk[ISS_QA_PER - 1] += 2.*c_tot*tot + n*c_tot*c_tot;
            blk[ISS_QA_PER - 2] = tot + n*c_tot;
        }
    }
}

// FSSW::computeAvgTotalEnergyMomentum (FSSW.cpp:2028-2059) from the QA block (spectators included)
void GpuFSSW::computeAvgTotalEnergyMomentum() {
    info("Averaged total energy and momentum:");
    for (int i = 0; i < 4; i++) {
        const double S1 = qa_[1 + i], S2 = qa_[5 + i];
        const double n = (qa_ranks_ > 1) ? qa_[0] : static_cast<double>(nev_);
        const double avg = S1/n;
        const double sq = S2/n;
        const double err = std::sqrt(std::max(0., sq - avg*avg)/n);
        std::ostringstream os;
        os << "<P[" << i << "]> = " << avg << " +/- " << err << " GeV.";
        info(os.str());
    }
}

void GpuFSSW::combine_samples_to_OSCAR() {
    iss_writers::write_oscar("OSCAR.DAT", table_path_ + "/OSCAR_header.txt", hadrons_,
                             event_off_.data(), nev_);
}

void GpuFSSW::combine_samples_to_gzip_file() {
    iss_writers::write_gzip("particle_samples.gz", hadrons_, event_off_.data(), nev_);
}

void GpuFSSW::combine_samples_to_binary_file() {
    iss_writers::write_binary("particle_samples.bin", hadrons_, event_off_.data(), nev_);
}
