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

#include "gpu_fssw.h"
#include "logger.h"
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
    clear();
    delete paraRdr_ptr;
}

void iSS::clear() {
    FOsurf_LRF_array_.clear();
    FOsurf_Tmunu_.clear();
    for (auto &p : particle_)
        for (auto *ch : p.decay_channels) delete ch;
    particle_.clear();
}

void iSS::require_fssw_() const {
    if (paraRdr_ptr->getVal("MC_sampling") != 4) {
        iss_host::error("the B200 engine implements the FSSW sampler only: set MC_sampling = 4 "
                        "(the legacy EmissionFunctionArray samplers MC_sampling = 1/2/3 are out of scope)");
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
    require_fssw_();
    std::vector<FO_surf> cells;
    read_FOdata reader(paraRdr_ptr, path_, table_path_, particle_table_path_);
    reader.read_in_freeze_out_data(cells, surface_filename_);
    info("total number of cells: " + std::to_string(cells.size()));
    if (cells.empty()) {
        iss_host::warning("No freeze-out fluid cell, exit now ...");
        exit(1);
    }
    afterburner_type_ = reader.get_afterburner_type();
    reader.read_in_chemical_potentials(cells, particle_);
    flag_PCE_ = reader.get_flag_PCE();
    computeFOSurfTmunu(cells);
    FOsurf_LRF_array_.clear();
    transform_to_local_rest_frame(cells, FOsurf_LRF_array_);
    info(" -- Read in data finished!");
    return 0;
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
    spectra_sampler_.reset(new GpuFSSW(randomSeed_, chosen, particle_, FOsurf_LRF_array_, flag_PCE_,
                                       paraRdr_ptr, path_, table_path_, afterburner_type_));
    return 0;
}

// iSS.cpp:130-165
int iSS::generate_samples() {
    info("Start computation and generating samples ...");
    prepare_sampler();
    spectra_sampler_->shell();
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
    info("Transforming fluid cells to their local rest frame ...");
    FOsurf_LRF_ptr.reserve(FOsurf_LRF_ptr.size() + FOsurf_ptr.size());
    for (const FO_surf &c : FOsurf_ptr) {
        const EtaRotation rot(c.eta);
        const float ut = rot.time_like(c.u0, c.u3);
        const float uz = rot.z_like(c.u0, c.u3);
        const float ux = c.u1, uy = c.u2;
        const double g = ut + 1.;
        const double L[4][4] = {{ut, -ux, -uy, -uz},
                                {-ux, 1. + ux*ux/g, ux*uy/g, ux*uz/g},
                                {-uy, ux*uy/g, 1. + uy*uy/g, uy*uz/g},
                                {-uz, ux*uz/g, uy*uz/g, 1. + uz*uz/g}};
        // contravariant surface normal in (t,x,y,z) from the Milne covariant components
        const Vec4 dsigma = {c.tau*c.da0*rot.ch - c.da3*rot.sh, -c.tau*c.da1, -c.tau*c.da2,
                             -c.da3*rot.ch + c.tau*c.da0*rot.sh};
        const Vec4 ds = boost_apply(L, dsigma);
        if (ds[0] < 0) continue;    // u.dsigma < 0: cell dropped (iSS.cpp:226)

        FO_surf_LRF o;
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
        FOsurf_LRF_ptr.push_back(o);
    }
}

// iSS.cpp:378-445: unweighted sum of the cells' T^{mu nu}; meaningful for one-cell inputs
// (the closure tests), kept because perform_checks prints it.
void iSS::computeFOSurfTmunu(std::vector<FO_surf> &FOsurf_ptr) {
    FOsurf_Tmunu_.assign(16, 0.f);
    FOsurf_Q_.assign(3, 0.f);
    for (const FO_surf &c : FOsurf_ptr) {
        const EtaRotation rot(c.eta);
        const float u[4] = {rot.time_like(c.u0, c.u3), c.u1, c.u2, rot.z_like(c.u0, c.u3)};
        float pi_tz[4][4];
        shear_to_tz(c, rot, pi_tz);
        for (int i = 0; i < 4; i++)
            for (int j = 0; j < 4; j++) {
                const float gij = (i != j) ? 0.f : (i == 0 ? 1.f : -1.f);
                const float Tij = (c.Edec*u[i]*u[j] - (c.Pdec + c.bulkPi)*(gij - u[i]*u[j])
                                   + pi_tz[i][j]);
                FOsurf_Tmunu_[4*i + j] += Tij;
            }
        FOsurf_Q_[0] = c.Bn;
        FOsurf_Q_[2] = 0.4*c.Bn;
    }
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
    const double nev = get_number_of_sampled_events();
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
    const double nev = get_number_of_sampled_events();
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
