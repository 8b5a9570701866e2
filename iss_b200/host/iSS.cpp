// iSS.cpp -- facade of the B200 Cooper-Frye engine; see iSS.h.
// Citations are to the reference's src/iSS.cpp.
#include "iSS.h"

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <sstream>

#include "engine_pool.h"
#include "gpu_fssw.h"
#include "gpu_spectra.h"
#include "logger.h"
#include "parallel.h"
#include "readindata.h"

using iSS_data::Vec4;
using iss_host::info;

namespace {

// Milne -> (t, z) rotation of one cell; float members on purpose (the reference stores
// cosh/sinh of the space-time rapidity in floats, iSS.cpp:189-190).
struct EtaRotation {
    float ch, sh;
    explicit EtaRotation(float eta) : ch(cosh(eta)), sh(sinh(eta)) {}
    float time_like(float a0, float a3) const { return a0*ch + a3*sh; }
    float z_like(float a0, float a3) const { return a3*ch + a0*sh; }
};

// symmetric pi^{mu nu} in (t,x,y,z) components from its Milne components (iSS.cpp:246-268,
// the same block again at :401-423); the `2.` literals make those terms double.
void shear_to_tz(const FO_surf &c, const EtaRotation &r, float out[4][4]) {
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
Vec4 boost_apply(const double L[4][4], const Vec4 &x) {
    Vec4 y = {0., 0., 0., 0.};
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) y[i] += L[i][j]*x[j];
    return y;
}

}  // namespace

iSS::iSS(std::string path, std::string table_path, std::string particle_table_path,
         std::string inputfile, std::string surface_filename)
    : path_(path), table_path_(table_path), particle_table_path_(particle_table_path),
      surface_filename_(surface_filename) {
    flag_PCE_ = 0;
    afterburner_type_ = AfterburnerType::UrQMD;
    randomSeed_ = -1;
    seed_set_ = false;
    paraRdr_ptr = new ParameterReader;
    paraRdr_ptr->readFromFile(inputfile);
}

iSS::~iSS() {
    spectra_sampler_.reset();
    efa_.reset();
    clear();
    delete paraRdr_ptr;
}

void iSS::drop_packed_lrf_() {
    if (lrf_packed_) {
        const int device = iss_pool::default_device();
        iss_handle *h = iss_pool::acquire_handle(device);
        iss_pool::PinnedBlock b;
        b.ptr = lrf_packed_;
        b.bytes = lrf_packed_bytes_;
        iss_pool::pinned_release(h, b);
        iss_pool::release_handle(device, h);
    }
    lrf_packed_ = nullptr;
    lrf_packed_bytes_ = 0;
    lrf_packed_n_ = -1;
}

void iSS::ensure_packed_lrf_() {
    const int64_t n = static_cast<int64_t>(FOsurf_LRF_array_.size());
    if (lrf_packed_ && lrf_packed_n_ == n) return;
    drop_packed_lrf_();
    if (n == 0) return;
    const int device = iss_pool::default_device();
    iss_handle *h = iss_pool::acquire_handle(device);
    iss_pool::PinnedBlock b = iss_pool::pinned_acquire(h, n*ISS_NFIELD*static_cast<int64_t>(sizeof(float)));
    iss_pool::release_handle(device, h);
    lrf_packed_ = b.ptr;
    lrf_packed_bytes_ = b.bytes;
    lrf_packed_n_ = n;
    GpuFSSW::pack_surface(FOsurf_LRF_array_, static_cast<float *>(lrf_packed_), 0, n);
}

void iSS::clear() {
    drop_packed_lrf_();
    FOsurf_array_.clear();
    FOsurf_LRF_array_.clear();
    FOsurf_Tmunu_.clear();
    for (auto &p : particle_)
        for (auto *ch : p.decay_channels) delete ch;
    particle_.clear();
}

void iSS::require_fssw_() const {
    const double mode = paraRdr_ptr->getVal("MC_sampling");
    if (mode != 4 && mode != 2) {
        iss_host::error("the B200 engine samples with FSSW (MC_sampling = 4) or with the legacy "
                        "conventional sampler (MC_sampling = 2); the legacy EmissionFunctionArray "
                        "samplers MC_sampling = 1/3 are out of scope");
        exit(-1);
    }
}

// MC_sampling = 4: FSSW sampler; MC_sampling = 2: the legacy "conventional" sampler of
// EmissionFunctionArray; MC_sampling = 0: smooth spectra and flows of the legacy class
// (calculate_vn = 1) without sampling.
void iSS::require_supported_mode_() const {
    const double mode = paraRdr_ptr->getVal("MC_sampling");
    if (mode != 4 && mode != 2 && mode != 0) {
        iss_host::error("the B200 engine implements MC_sampling = 4 (FSSW), MC_sampling = 2 (legacy "
                        "conventional sampler) and MC_sampling = 0 (smooth spectra and flows); the "
                        "legacy EmissionFunctionArray samplers MC_sampling = 1/3 are out of scope");
        exit(-1);
    }
}

int iSS::shell() {
    if (read_in_FO_surface() != 0) {
        iss_host::error("Some errors happened in reading in the hyper-surface");
        exit(-1);
    }
    set_random_seed();
    if (generate_samples() != 0) {
        iss_host::error("Some errors happened in generating particle samples");
        exit(-1);
    }
    return 0;
}

// iSS.cpp:86-113
int iSS::read_in_FO_surface() {
    require_supported_mode_();
    const bool fssw = paraRdr_ptr->getVal("MC_sampling") == 4;
    const char *prof_env = getenv("ISS_PROFILE");
    const bool prof = prof_env && atoi(prof_env) == 1;
    auto t0 = std::chrono::steady_clock::now();
    auto lap = [&](const char *what) {
        const auto t1 = std::chrono::steady_clock::now();
        if (prof)
            fprintf(stderr, "[iss profile] %-28s %8.3f ms\n", what,
                    1e3*std::chrono::duration<double>(t1 - t0).count());
        t0 = t1;
    };
    std::vector<FO_surf> cells;
    read_FOdata reader(paraRdr_ptr, path_, table_path_, particle_table_path_);
    lap("reader ctor (EOS table)");
    lrf_packed_n_ = -1;         // a new surface: the packed copy is rebuilt
    FOsurf_LRF_array_.clear();
    FOsurf_array_.clear();
    // (the blocked binary pipeline feeds the LRF transform; the lab-frame path keeps whole cells)
    const int64_t nbin = fssw ? reader.open_binary_surface(surface_filename_) : -1;
    // ISS_INGEST=host keeps the per-cell work on the host (tests without a GPU, debugging); the
    // default for binary surfaces is the device (no silent fallback: without a CUDA device
    // acquire_handle exits with a message)
    const char *ingest_env = getenv("ISS_INGEST");
    const bool device_ingest = !(ingest_env && std::string(ingest_env) == "host");
    if (nbin >= 0 && device_ingest) {
        std::cout << " -- Read spatial positions of freeze out surface from MUSIC...";
        ingest_binary_on_device_(reader, nbin);
        reader.close_surface();
        std::cout << "done" << std::endl;
        lap("binary surface, device ingest");
        std::vector<FO_surf> none;
        afterburner_type_ = reader.get_afterburner_type();
        reader.read_in_chemical_potentials(none, particle_);
        flag_PCE_ = reader.get_flag_PCE();
        report_Tmunu_();
        lap("particle table");
        ensure_packed_lrf_();
        lap("pack LRF surface (pinned)");
    } else if (nbin >= 0) {
        // binary surface: parse -> regulate -> T^{mu nu} -> LRF transform over blocks that stay in
        // cache; per-cell arithmetic and cell order are those of the whole-surface path
        std::cout << " -- Read spatial positions of freeze out surface from MUSIC...";
        FOsurf_Tmunu_.assign(16, 0.f);
        FOsurf_Q_.assign(3, 0.f);
        FOsurf_LRF_array_.reserve(nbin);
        const int64_t BLOCK = 1 << 17;
        int64_t ntotal = 0;
        double tp[4] = {0, 0, 0, 0};
        auto now = [] { return std::chrono::steady_clock::now(); };
        auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
            return 1e3*std::chrono::duration<double>(b - a).count();
        };
        lap("open (slurp)");
        for (int64_t c0 = 0; c0 < nbin; c0 += BLOCK) {
            auto a0 = now();
            reader.read_binary_block(cells, c0, std::min<int64_t>(BLOCK, nbin - c0));
            auto a1 = now();
            reader.regulate_surface_cells(cells, c0 == 0);
            auto a2 = now();
            ntotal += static_cast<int64_t>(cells.size());
            accumulate_Tmunu_(cells);
            auto a3 = now();
            transform_to_local_rest_frame(cells, FOsurf_LRF_array_);
            auto a4 = now();
            tp[0] += ms(a0, a1); tp[1] += ms(a1, a2); tp[2] += ms(a2, a3); tp[3] += ms(a3, a4);
        }
        if (prof)
            fprintf(stderr, "[iss profile]   parse %.1f regulate %.1f tmunu %.1f lrf %.1f ms\n", tp[0],
                    tp[1], tp[2], tp[3]);
        reader.close_surface();
        std::cout << "done" << std::endl;
        lap("binary surface pipeline");
        info("total number of cells: " + std::to_string(ntotal));
        if (ntotal == 0) {
            iss_host::warning("No freeze-out fluid cell, exit now ...");
            exit(1);
        }
        cells.clear();
        afterburner_type_ = reader.get_afterburner_type();
        reader.read_in_chemical_potentials(cells, particle_);
        flag_PCE_ = reader.get_flag_PCE();
        report_Tmunu_();
        lap("particle table");
    } else {
        reader.read_in_freeze_out_data(cells, surface_filename_);
        lap("read + regulate surface");
        info("total number of cells: " + std::to_string(cells.size()));
        if (cells.empty()) {
            iss_host::warning("No freeze-out fluid cell, exit now ...");
            exit(1);
        }
        afterburner_type_ = reader.get_afterburner_type();
        reader.read_in_chemical_potentials(cells, particle_);
        flag_PCE_ = reader.get_flag_PCE();
        lap("particle table");
        computeFOSurfTmunu(cells);
        lap("computeFOSurfTmunu");
        if (fssw) {
            transform_to_local_rest_frame(cells, FOsurf_LRF_array_);
            lap("transform_to_local_rest_frame");
        } else {
            FOsurf_array_.swap(cells);      // iSS.cpp:105-109
        }
    }
    info(" -- Read in data finished!");
    return 0;
}

// Binary surface -> FOsurf_LRF_array_ with the per-cell work on the GPU
// (iss_cuda_ingest_music_binary, include/iss_cuda.h): the file is sent in chunks, the compacted
// local-rest-frame records and per-cell T^{mu nu} tensors come back; what stays on the host is
// the sequential float sum of the tensors in file order (iSS.cpp:378-445) and the messages.
void iSS::ingest_binary_on_device_(read_FOdata &reader, int64_t nbin) {
    const int device = iss_pool::default_device();
    iss_handle *h = iss_pool::acquire_handle(device);
    iss_ingest_options opt;
    opt.boost_invariant = reader.boost_invariant() ? 1 : 0;
    opt.regulate_eos = reader.regulates_eos() ? 1 : 0;
    opt.hrg_nB = reader.hrg_nB();
    opt.reserved = 0;
    opt.hrg_rows = static_cast<int64_t>(reader.hrg_table().size()/7);
    FOsurf_Tmunu_.assign(16, 0.f);
    FOsurf_Q_.assign(3, 0.f);
    const int64_t CHUNK = 1 << 22;          // cells per call: bounds the device arena (~2 GB)
    const int64_t cap = std::min<int64_t>(nbin, CHUNK);
    iss_pool::PinnedBlock blk = iss_pool::pinned_acquire(
        h, cap*static_cast<int64_t>((ISS_NFIELD + 16)*sizeof(float) + 1));
    float *lrf = static_cast<float *>(blk.ptr);
    float *tm = lrf + cap*ISS_NFIELD;
    uint8_t *status = reinterpret_cast<uint8_t *>(tm + cap*16);
    const float *raw = reader.binary_records();
    const bool regulate = reader.regulates_eos();
    if (regulate) std::cout << "Regulate local temperature with pure HRG EoS." << std::endl;
    int64_t ntotal = 0;
    float last_Bn = 0.f;
    bool any_T = false;
    for (int64_t c0 = 0; c0 < nbin; c0 += CHUNK) {
        const int64_t n = std::min<int64_t>(CHUNK, nbin - c0);
        iss_ingest_result res;
        const int rc = iss_cuda_ingest_music_binary(h, raw + 34*c0, n, &opt, reader.hrg_table().data(),
                                                    lrf, tm, status, &res);
        if (rc != ISS_OK) {
            iss_host::error(std::string("iss_cuda_ingest_music_binary failed: ") + iss_cuda_last_error(h));
            exit(-1);
        }
        // the reference's per-cell messages (readindata.cpp:752-758, 1262-1266)
        if (res.n_after_T != n || regulate) {
            for (int64_t c = 0; c < n; c++) {
                if (status[c] & ISS_INGEST_DROPPED_T) {
                    const float *a = raw + 34*(c0 + c);
                    std::cout << "Discard surf elem: T = " << static_cast<float>(a[13]*iSS_data::hbarC)
                              << " GeV, Edec = " << static_cast<float>(a[12]*iSS_data::hbarC)
                              << " GeV/fm^3, rhoB = " << a[29] << " 1/fm^3, muB = "
                              << static_cast<float>(a[14]*iSS_data::hbarC) << " GeV. " << std::endl;
                } else if (status[c] & ISS_INGEST_EOS_RANGE) {
                    std::ostringstream os;
                    os << "ed is out of range: ed = "
                       << static_cast<double>(static_cast<float>(raw[34*(c0 + c) + 12]*iSS_data::hbarC))
                       << " GeV/fm^3. Can not regulate this fluid cell!";
                    iss_host::warning(os.str());
                }
            }
        }
        // sequential float sums of the per-cell tensors, file order
        float acc[16];
        for (int k = 0; k < 16; k++) acc[k] = FOsurf_Tmunu_[k];
        for (int64_t ic = 0; ic < res.n_after_T; ic++)
            for (int k = 0; k < 16; k++) acc[k] += tm[ic*16 + k];
        for (int k = 0; k < 16; k++) FOsurf_Tmunu_[k] = acc[k];
        ntotal += res.n_after_T;
        // Bn of the last cell that passed the T filter (FOsurf_Q_, iSS.cpp:441-444)
        for (int64_t c = n - 1; c >= 0; c--)
            if (!(status[c] & ISS_INGEST_DROPPED_T)) {
                last_Bn = raw[34*(c0 + c) + 29];
                any_T = true;
                break;
            }
        // records -> FO_surf_LRF
        const size_t base = FOsurf_LRF_array_.size();
        FOsurf_LRF_array_.resize(base + res.n_kept);
        iss_host::parallel_ranges(res.n_kept, iss_host::ingest_threads(res.n_kept),
                                  [&](int64_t b, int64_t e, int) {
            for (int64_t i = b; i < e; i++) {
                const float *r = lrf + i*ISS_NFIELD;
                FO_surf_LRF &o = FOsurf_LRF_array_[base + i];
                o.tau = r[ISS_F_TAU]; o.xpt = r[ISS_F_X]; o.ypt = r[ISS_F_Y]; o.eta = r[ISS_F_ETA];
                o.da_mu_LRF = {r[ISS_F_DA0], r[ISS_F_DA1], r[ISS_F_DA2], r[ISS_F_DA3]};
                o.u_tz = {r[ISS_F_UT], r[ISS_F_UX], r[ISS_F_UY], r[ISS_F_UZ]};
                o.Edec = r[ISS_F_E]; o.Tdec = r[ISS_F_T]; o.Pdec = r[ISS_F_P]; o.Bn = r[ISS_F_NB];
                o.muB = r[ISS_F_MUB]; o.muS = r[ISS_F_MUS]; o.muQ = r[ISS_F_MUQ];
                o.bulkPi = r[ISS_F_BULKPI];
                o.piLRF_xx = r[ISS_F_PIXX]; o.piLRF_xy = r[ISS_F_PIXY]; o.piLRF_xz = r[ISS_F_PIXZ];
                o.piLRF_yy = r[ISS_F_PIYY]; o.piLRF_yz = r[ISS_F_PIYZ];
                o.qmuLRF_x = r[ISS_F_QX]; o.qmuLRF_y = r[ISS_F_QY]; o.qmuLRF_z = r[ISS_F_QZ];
            }
        });
    }
    if (any_T) {
        FOsurf_Q_[0] = last_Bn;
        FOsurf_Q_[2] = 0.4*last_Bn;
    }
    iss_pool::pinned_release(h, blk);
    iss_pool::release_handle(device, h);
    info("total number of cells: " + std::to_string(ntotal));
    if (ntotal == 0) {
        iss_host::warning("No freeze-out fluid cell, exit now ...");
        exit(1);
    }
}

// Philox needs a definite key: a negative seed (reference: std::random_device, Random.cpp:7-14)
// is replaced by one draw from the OS entropy source.
void iSS::set_random_seed() { set_random_seed(static_cast<int>(paraRdr_ptr->getVal("randomSeed"))); }

void iSS::set_random_seed(int randomSeed_in) {
    randomSeed_ = randomSeed_in;
    if (randomSeed_ < 0) {
        unsigned int v = 0;
        FILE *f = fopen("/dev/urandom", "rb");
        if (f) {
            if (fread(&v, sizeof(v), 1, f) != 1) v = 0;
            fclose(f);
        }
        randomSeed_ = static_cast<long>(v >> 1);
    }
    seed_set_ = true;
}

// chosen_particles_*.dat: one Monte-Carlo id per line (iSS.cpp:132-142)
std::vector<int> iSS::read_chosen_particles() const {
    std::string list = particle_table_path_;
    if (afterburner_type_ == AfterburnerType::SMASH) list += "/chosen_particles_SMASH.dat";
    else if (afterburner_type_ == AfterburnerType::UrQMD) list += "/chosen_particles_urqmd_v3.3+.dat";
    else list += "/chosen_particles_s95p-v1.dat";
    std::ifstream in(list.c_str());
    if (!in.good()) {
        iss_host::error("Can not found file: " + list);
        exit(-1);
    }
    std::vector<int> chosen;
    double v;
    while (in >> v) chosen.push_back(static_cast<int>(v));
    return chosen;
}

int iSS::prepare_sampler() {
    const auto t0 = std::chrono::steady_clock::now();
    struct Report {
        std::chrono::steady_clock::time_point t0;
        ~Report() {
            const char *e = getenv("ISS_PROFILE");
            if (e && atoi(e) == 1)
                fprintf(stderr, "[iss profile] %-28s %8.3f ms\n", "prepare_sampler total",
                        1e3*std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
        }
    } report{t0};
    require_fssw_();
    if (!seed_set_) set_random_seed();
    const std::vector<int> chosen = read_chosen_particles();
    spectra_sampler_.reset();   // frees the previous batch before the new one is allocated
    if (paraRdr_ptr->getVal("MC_sampling") == 2) {
        // legacy conventional sampler over the lab-frame cells (iSS.cpp:151-163)
        spectra_sampler_.reset(new GpuFSSW(randomSeed_, chosen, particle_, FOsurf_array_, flag_PCE_,
                                           paraRdr_ptr, path_, table_path_, afterburner_type_));
        return 0;
    }
    ensure_packed_lrf_();
    spectra_sampler_.reset(new GpuFSSW(randomSeed_, chosen, particle_, FOsurf_LRF_array_, flag_PCE_,
                                       paraRdr_ptr, path_, table_path_, afterburner_type_,
                                       static_cast<const float *>(lrf_packed_)));
    return 0;
}

// iSS.cpp:130-165
int iSS::generate_samples() {
    info("Start computation and generating samples ...");
    require_supported_mode_();
    const double mc_mode = paraRdr_ptr->getVal("MC_sampling");
    if (mc_mode == 2 && paraRdr_ptr->getVal("calculate_vn") == 1) {
        // the legacy class writes the smooth spectra and flows before it samples
        // (EmissionFunctionArray::shell, emissionfunction.cpp:2554-2561)
        const std::vector<int> chosen = read_chosen_particles();
        efa_.reset();
        efa_.reset(new GpuSpectra(chosen, particle_, FOsurf_array_, flag_PCE_, paraRdr_ptr, path_,
                                  table_path_, afterburner_type_));
        efa_->shell();
    }
    if (mc_mode == 4 || mc_mode == 2) {
        prepare_sampler();
        spectra_sampler_->shell();
    } else {
        const std::vector<int> chosen = read_chosen_particles();
        efa_.reset();
        efa_.reset(new GpuSpectra(chosen, particle_, FOsurf_array_, flag_PCE_, paraRdr_ptr, path_,
                                  table_path_, afterburner_type_));
        efa_->shell();
    }
    return 0;
}

int iSS::get_number_of_sampled_events() {
    return spectra_sampler_ ? spectra_sampler_->get_number_of_sampled_events() : 0;
}

int iSS::get_number_of_particles(int iev) { return spectra_sampler_->get_number_of_particles(iev); }

iSS_Hadron iSS::get_hadron(int iev, int ipart) { return spectra_sampler_->get_hadron(iev, ipart); }

std::vector<iSS_Hadron> *iSS::get_hadron_list_iev(const int iev) {
    return spectra_sampler_->get_hadron_list_iev(iev);
}

// iSS.cpp:170-293.  Mixed float/double arithmetic is part of the contract: the yields are
// compared with the reference at 1e-6 and Sigma_LRF = da_mu_LRF[0] multiplies every yield.
void iSS::transform_to_local_rest_frame(std::vector<FO_surf> &FOsurf_ptr,
                                        std::vector<FO_surf_LRF> &FOsurf_LRF_ptr) {
    const int64_t ncell = static_cast<int64_t>(FOsurf_ptr.size());
    const int nthread = iss_host::ingest_threads(ncell);
    // pass 1: which cells survive u.dsigma >= 0 (iSS.cpp:226), counted per thread range;
    // pass 2: full transform written straight to the final position (file order is kept)
    auto boost_of = [](const FO_surf &c, const EtaRotation &rot, double L[4][4], float &ut, float &uz) {
        ut = rot.time_like(c.u0, c.u3);
        uz = rot.z_like(c.u0, c.u3);
        const float ux = c.u1, uy = c.u2;
        const double g = ut + 1.;
        const double M[4][4] = {{ut, -ux, -uy, -uz},
                                {-ux, 1. + ux*ux/g, ux*uy/g, ux*uz/g},
                                {-uy, ux*uy/g, 1. + uy*uy/g, uy*uz/g},
                                {-uz, ux*uz/g, uy*uz/g, 1. + uz*uz/g}};
        for (int i = 0; i < 4; i++)
            for (int j = 0; j < 4; j++) L[i][j] = M[i][j];
    };
    auto normal_of = [](const FO_surf &c, const EtaRotation &rot) {
        // contravariant surface normal in (t,x,y,z) from the Milne covariant components
        const Vec4 dsigma = {c.tau*c.da0*rot.ch - c.da3*rot.sh, -c.tau*c.da1, -c.tau*c.da2,
                             -c.da3*rot.ch + c.tau*c.da0*rot.sh};
        return dsigma;
    };
    std::vector<unsigned char> keep(static_cast<size_t>(ncell));
    std::vector<int64_t> count(nthread + 1, 0);
    iss_host::parallel_ranges(ncell, nthread, [&](int64_t c0, int64_t c1, int t) {
        int64_t n = 0;
        for (int64_t ic = c0; ic < c1; ic++) {
            const FO_surf &c = FOsurf_ptr[ic];
            const EtaRotation rot(c.eta);
            double L[4][4];
            float ut, uz;
            boost_of(c, rot, L, ut, uz);
            const Vec4 ds = boost_apply(L, normal_of(c, rot));
            keep[ic] = !(ds[0] < 0);
            n += keep[ic];
        }
        count[t + 1] = n;
    });
    for (int t = 0; t < nthread; t++) count[t + 1] += count[t];
    const size_t base = FOsurf_LRF_ptr.size();
    FOsurf_LRF_ptr.resize(base + count[nthread]);
    iss_host::parallel_ranges(ncell, nthread, [&](int64_t c0, int64_t c1, int t) {
        size_t w = base + count[t];
        for (int64_t ic = c0; ic < c1; ic++) {
            if (!keep[ic]) continue;
            const FO_surf &c = FOsurf_ptr[ic];
            const EtaRotation rot(c.eta);
            double L[4][4];
            float ut, uz;
            boost_of(c, rot, L, ut, uz);
            const float ux = c.u1, uy = c.u2;
            const Vec4 ds = boost_apply(L, normal_of(c, rot));

            FO_surf_LRF &o = FOsurf_LRF_ptr[w++];
            o.tau = c.tau; o.xpt = c.xpt; o.ypt = c.ypt; o.eta = c.eta;
            o.Edec = c.Edec; o.Tdec = c.Tdec; o.Pdec = c.Pdec;
            o.Bn = c.Bn; o.muB = c.muB; o.muS = c.muS; o.muQ = c.muQ;
            o.bulkPi = c.bulkPi;
            o.particle_mu_PCE = c.particle_mu_PCE;
            o.u_tz = {ut, ux, uy, uz};
            o.da_mu_LRF = {ds[0], -ds[1], -ds[2], -ds[3]};

            const Vec4 q_tz = {rot.time_like(c.qmu0, c.qmu3), c.qmu1, c.qmu2,
                               rot.z_like(c.qmu0, c.qmu3)};
            const Vec4 q = boost_apply(L, q_tz);
            o.qmuLRF_x = q[1]; o.qmuLRF_y = q[2]; o.qmuLRF_z = q[3];

            float pi_tz[4][4];
            shear_to_tz(c, rot, pi_tz);
            float pi_lrf[4][4];
            for (int i = 1; i < 3; i++)         // only xx, xy, xz, yy, yz are kept
                for (int j = i; j < 4; j++) {
                    float acc = 0.;
                    for (int a = 0; a < 4; a++)
                        for (int b = 0; b < 4; b++) acc += (L[i][a]*pi_tz[a][b]*L[b][j]);
                    pi_lrf[i][j] = acc;
                }
            o.piLRF_xx = pi_lrf[1][1]; o.piLRF_xy = pi_lrf[1][2]; o.piLRF_xz = pi_lrf[1][3];
            o.piLRF_yy = pi_lrf[2][2]; o.piLRF_yz = pi_lrf[2][3];
        }
    });
}

// iSS.cpp:378-445: unweighted sum of the cells' T^{mu nu}; meaningful for one-cell inputs
// (the closure tests), kept because perform_checks prints it.
void iSS::computeFOSurfTmunu(std::vector<FO_surf> &FOsurf_ptr) {
    FOsurf_Tmunu_.assign(16, 0.f);
    FOsurf_Q_.assign(3, 0.f);
    accumulate_Tmunu_(FOsurf_ptr);
    report_Tmunu_();
}

void iSS::accumulate_Tmunu_(const std::vector<FO_surf> &FOsurf_ptr) {
    // the per-cell tensors are computed by several threads, block by block; the accumulation
    // stays a sequential float sum in file order like the reference's
    const int64_t ncell = static_cast<int64_t>(FOsurf_ptr.size());
    const int64_t BLOCK = 1 << 17;
    std::vector<float> tmp(static_cast<size_t>(std::min<int64_t>(BLOCK, std::max<int64_t>(ncell, 1)))*16);
    for (int64_t b0 = 0; b0 < ncell; b0 += BLOCK) {
        const int64_t nb = std::min<int64_t>(BLOCK, ncell - b0);
        iss_host::parallel_ranges(nb, iss_host::ingest_threads(nb), [&](int64_t c0, int64_t c1, int) {
            for (int64_t ic = c0; ic < c1; ic++) {
                const FO_surf &c = FOsurf_ptr[b0 + ic];
                const EtaRotation rot(c.eta);
                const float u[4] = {rot.time_like(c.u0, c.u3), c.u1, c.u2, rot.z_like(c.u0, c.u3)};
                float pi_tz[4][4];
                shear_to_tz(c, rot, pi_tz);
                for (int i = 0; i < 4; i++)
                    for (int j = 0; j < 4; j++) {
                        const float gij = (i != j) ? 0.f : (i == 0 ? 1.f : -1.f);
                        tmp[ic*16 + 4*i + j] = (c.Edec*u[i]*u[j]
                                                - (c.Pdec + c.bulkPi)*(gij - u[i]*u[j]) + pi_tz[i][j]);
                    }
            }
        });
        float acc[16];
        for (int k = 0; k < 16; k++) acc[k] = FOsurf_Tmunu_[k];
        const float *t = tmp.data();
        for (int64_t ic = 0; ic < nb; ic++)
            for (int k = 0; k < 16; k++) acc[k] += t[ic*16 + k];    // 16 independent running sums
        for (int k = 0; k < 16; k++) FOsurf_Tmunu_[k] = acc[k];
    }
    if (ncell > 0) {
        const FO_surf &c = FOsurf_ptr[ncell - 1];
        FOsurf_Q_[0] = c.Bn;
        FOsurf_Q_[2] = 0.4*c.Bn;
    }
}

void iSS::report_Tmunu_() const {
    info("The total energy-momentum tensor from the surface:");
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) {
            std::ostringstream os;
            os << "T[" << i << "][" << j << "] = " << std::scientific << std::setprecision(6)
               << FOsurf_Tmunu_[4*i + j] << " GeV/fm^3.";
            info(os.str());
        }
}

void iSS::getParticleQuantumNumbers(long monval, std::array<int, 3> &Qarr) {
    for (const auto &p : particle_)
        if (p.monval == monval) {
            Qarr = {p.baryon, p.strange, p.charge};
            return;
        }
}

// iSS.cpp:296-363, from the QA block the device accumulated while sampling
// (layout: include/iss_cuda.h).  Same file format as the reference.
void iSS::construct_Tmunu_from_particle_samples() {
    info("Constructing the fluid cell T^{mu nu} from samples ...");
    const std::vector<double> &qa = spectra_sampler_->qa_block();
    const double volume = FOsurf_LRF_array_[0].da_mu_LRF[0]/FOsurf_LRF_array_[0].u_tz[0];
    // the events behind the block: all ranks' when it was reduced (reduce_checks_over_ranks = 1)
    const double nev = (spectra_sampler_->qa_ranks() > 1) ? spectra_sampler_->qa_events()
                                                          : get_number_of_sampled_events();
    std::ofstream output("checkReconstructedTmunu.dat");
    output << "# Tmunu_FOcell[GeV/fm^3]  Tmunu_Particles[GeV/fm^3]  diff" << std::endl;
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) {
            const double T = qa[9 + 4*i + j]/(nev*volume);
            const double ref = FOsurf_Tmunu_[4*i + j];
            output << std::scientific << std::setprecision(8) << ref << "  " << T << "  "
                   << ref - T << std::endl;
            std::ostringstream os;
            os << "check: T[" << i << "][" << j << "] = " << ref << " GeV/fm^3,  " << T
               << " GeV/fm^3, diff = " << ref - T << " GeV/fm^3";
            info(os.str());
        }
    for (int i = 0; i < 3; i++) {
        const double n = qa[26 + i]/(nev*volume);
        if (i == 0)
            output << std::scientific << std::setprecision(8) << FOsurf_Q_[i] << "  " << n << "  "
                   << FOsurf_Q_[i] - n << std::endl;
        std::ostringstream os;
        os << "check: nQ[" << i << "] = " << FOsurf_Q_[i] << " 1/fm^3," << n
           << " 1/fm^3, diff = " << FOsurf_Q_[i] - n << " 1/fm^3";
        info(os.str());
    }
}

// iSS.cpp:59-83 + Histogram.cpp:39-61: pi+ and proton pT spectra, same columns as
// Histogram::output_histogram (x = mean pT of the bin, y = counts/event, error, counts).
void iSS::perform_checks() {
    info("Performing checks for the samples ...");
    construct_Tmunu_from_particle_samples();
    const std::vector<double> &qa = spectra_sampler_->qa_block();
    const double nev = (spectra_sampler_->qa_ranks() > 1) ? spectra_sampler_->qa_events()
                                                          : get_number_of_sampled_events();
    const char *files[2] = {"check_211_spectra.dat", "check_2212_spectra.dat"};
    const double bin_width = 5.0/(ISS_QA_NPT - 1);
    for (int k = 0; k < 2; k++) {
        const double *blk = qa.data() + ISS_QA_HEAD + static_cast<size_t>(k)*ISS_QA_PER;
        std::ofstream of(files[k]);
        of << "# x  y  y_err  bin_counts" << std::endl;
        for (int i = 0; i < ISS_QA_NPT; i++) {
            const double cnt = blk[i], sum = blk[ISS_QA_NPT + i], sq = blk[2*ISS_QA_NPT + i];
            const double x = (cnt > 0) ? sum/cnt : (i + 0.5)*bin_width;
            const double y_err = std::sqrt(sq/nev - cnt*cnt/(nev*nev))/std::sqrt(nev);
            of << std::scientific << std::setprecision(6) << std::setw(10) << x << "  " << cnt/nev
               << "  " << y_err << "  " << static_cast<long>(cnt) << std::endl;
        }
    }
}
