// gpu_spectra.cpp -- see gpu_spectra.h.  Citations are to the reference's src/emissionfunction.cpp.
#include "gpu_spectra.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <sstream>

#include "engine_pool.h"
#include "logger.h"

using iss_host::info;

namespace {

// two-column bin table (value, weight)
void load_bin_table(const std::string &file, std::vector<double> &x, std::vector<double> &w) {
    const std::vector<double> &v = iss_pool::cached_numbers(file, 0);
    if (v.empty() || v.size() % 2 != 0) {
        iss_host::error("bin table is not a two-column table: " + file);
        exit(-1);
    }
    x.resize(v.size()/2);
    w.resize(v.size()/2);
    for (size_t i = 0; i < x.size(); i++) {
        x[i] = v[2*i];
        w[i] = v[2*i + 1];
    }
}

}  // namespace

void GpuSpectra::check_(int rc, const char *what) {
    if (rc == ISS_OK) return;
    std::ostringstream os;
    os << what << " failed (status " << rc << "): " << (h_ ? iss_cuda_last_error(h_) : "no handle");
    iss_host::error(os.str());
    exit(-1);
}

GpuSpectra::GpuSpectra(const std::vector<int> &chosen_monvals,
                       const std::vector<particle_info> &particles,
                       const std::vector<FO_surf> &FOsurf, int flag_PCE, ParameterReader *paraRdr,
                       std::string path, std::string table_path, AfterburnerType)
    : paraRdr_(paraRdr), path_(path), table_path_(table_path), particles_(particles), surf_(FOsurf) {
    if (flag_PCE != 0) {
        iss_host::error("partial chemical equilibrium EoS is not supported by the B200 engine");
        exit(1);
    }
    // keys of the legacy class used on this path (emissionfunction.cpp:88-100)
    include_shear_ = static_cast<int>(paraRdr_->getVal("include_deltaf_shear"));
    include_bulk_ = static_cast<int>(paraRdr_->getVal("include_deltaf_bulk"));
    bulk_kind_ = static_cast<int>(paraRdr_->getVal("bulk_deltaf_kind"));
    include_diff_ = static_cast<int>(paraRdr_->getVal("include_deltaf_diffusion"));
    restrict_deltaf_ = static_cast<int>(paraRdr_->getVal("restrict_deltaf"));
    deltaf_max_ratio_ = paraRdr_->getVal("deltaf_max_ratio");
    use_pos_dN_only_ = static_cast<int>(paraRdr_->getVal("use_pos_dN_only"));
    grouping_particles_ = static_cast<int>(paraRdr_->getVal("grouping_particles"));
    grouping_tolerance_ = paraRdr_->getVal("grouping_tolerance");
    MC_sampling_ = static_cast<int>(paraRdr_->getVal("MC_sampling"));

    load_bin_table(table_path_ + "/bin_tables/pT_gauss_table.dat", pT_, pT_w_);
    load_bin_table(table_path_ + "/bin_tables/phi_gauss_table.dat", phi_, phi_w_);
    load_bin_table(table_path_ + "/bin_tables/eta_uni_table.dat", eta_, eta_w_);
    npT_ = static_cast<int>(pT_.size());
    nphi_ = static_cast<int>(phi_.size());
    neta_ = static_cast<int>(eta_.size());

    // chosen particles (emissionfunction.cpp:144-207): a 0/1 table over the pdg list for the
    // historic output, and the list of indices in file order, mass-sorted when grouping is on
    const int Nparticles = static_cast<int>(particles_.size());
    chosen_particles_01_table_.assign(Nparticles, 0);
    std::vector<int> missing;
    for (int monval : chosen_monvals) {
        int found = -1;
        for (int n = 0; n < Nparticles; n++)
            if (particles_[n].monval == monval) {
                found = n;
                break;
            }
        if (found < 0) {
            missing.push_back(monval);
            continue;
        }
        chosen_particles_01_table_[found] = 1;
        chosen_particles_sampling_table_.push_back(found);
    }
    if (!missing.empty()) {
        iss_host::warning("not all chosen particles are in the pdg particle list!");
        iss_host::warning("There are " + std::to_string(missing.size())
                          + " particles can not be found in the pdg particle list!");
        iss_host::warning("Their monte carlo numbers are:");
        for (int m : missing) iss_host::warning(std::to_string(m));
    }
    if (grouping_particles_)    // the reference's bubble sort swaps on strict >, i.e. it is stable
        std::stable_sort(chosen_particles_sampling_table_.begin(),
                         chosen_particles_sampling_table_.end(),
                         [&](int a, int b) { return particles_[a].mass < particles_[b].mass; });
    dN_.assign(Nparticles, std::vector<double>());

    device_ = iss_pool::default_device();
    h_ = iss_pool::acquire_handle(device_);
    upload_surface_();
    upload_tables_();
}

GpuSpectra::~GpuSpectra() {
    if (!h_) return;
    iss_cuda_synchronize(h_);
    iss_pool::release_handle(device_, h_);
}

// FO_surf -> the ISS_L_* record of include/iss_cuda.h
void GpuSpectra::upload_surface_() {
    const int64_t n = static_cast<int64_t>(surf_.size());
    std::vector<float> rec(static_cast<size_t>(n)*ISS_LAB_NFIELD);
    for (int64_t l = 0; l < n; l++) {
        const FO_surf &c = surf_[l];
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
    }
    check_(iss_cuda_upload_surface_lab(h_, rec.data(), n), "iss_cuda_upload_surface_lab");
}

// diffusion coefficient table (load_deltaf_qmu_coeff_table, emissionfunction.cpp:3764-3786):
// 100 (mu_B) x 150 (T) rows "T muB kappa", T fastest, hard-coded grid
void GpuSpectra::upload_tables_() {
    if (include_diff_ != 1) return;
    const std::string file = table_path_ + "/deltaf_tables/Coefficients_RTA_diffusion.dat";
    const std::vector<double> &v = iss_pool::cached_numbers(file, 0);
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

bool GpuSpectra::particles_are_the_same(int idx1, int idx2) const {
    const particle_info &a = particles_[idx1], &b = particles_[idx2];
    if (a.sign != b.sign || a.gspin != b.gspin || a.baryon != b.baryon || a.strange != b.strange
        || a.charge != b.charge)
        return false;
    return !(std::abs((a.mass - b.mass)/(b.mass + 1e-30)) > grouping_tolerance_);
}

void GpuSpectra::compute_tables_(const std::vector<int> &particle_idx) {
    if (particle_idx.empty()) return;
    std::vector<iss_species> sp(particle_idx.size());
    for (size_t k = 0; k < sp.size(); k++) {
        const particle_info &p = particles_[particle_idx[k]];
        memset(&sp[k], 0, sizeof(iss_species));
        sp[k].pid = p.monval;
        sp[k].gspin = p.gspin;
        sp[k].baryon = p.baryon;
        sp[k].strange = p.strange;
        sp[k].charge = p.charge;
        sp[k].sign = p.sign;
        sp[k].decay_idx = -1;
        sp[k].mass = p.mass;
    }
    iss_spectra_options o;
    memset(&o, 0, sizeof(o));
    o.include_deltaf_shear = include_shear_;
    o.include_deltaf_bulk = include_bulk_;
    o.bulk_deltaf_kind = bulk_kind_;
    o.include_deltaf_diffusion = include_diff_;
    o.restrict_deltaf = restrict_deltaf_;
    o.use_pos_dN_only = use_pos_dN_only_;
    o.deltaf_max_ratio = deltaf_max_ratio_;
    const size_t npt = static_cast<size_t>(npT_)*nphi_;
    std::vector<double> out(sp.size()*npt);
    check_(iss_cuda_spectra(h_, &o, sp.data(), static_cast<int32_t>(sp.size()), pT_.data(), npT_,
                            phi_.data(), nphi_, eta_.data(), eta_w_.data(), neta_, out.data(), nullptr),
           "iss_cuda_spectra");
    double ev = 0., ms = 0.;
    iss_cuda_spectra_stats(h_, &ev, &ms);
    evaluations_ += ev;
    kernel_ms_ += ms;
    for (size_t k = 0; k < sp.size(); k++)
        dN_[particle_idx[k]].assign(out.begin() + k*npt, out.begin() + (k + 1)*npt);
}

void GpuSpectra::print_dN_table_(FILE *f, const std::vector<double> *dN) const {
    for (int j = 0; j < nphi_; j++) {
        for (int i = 0; i < npT_; i++)
            fprintf(f, "%15.8e  ", dN ? (*dN)[static_cast<size_t>(i)*nphi_ + j] : 0.0);
        fputc('\n', f);
    }
}

void GpuSpectra::calculate_flows_(const std::vector<double> &dN, double mass, int to_order,
                                  const std::string &flow_differential_filename,
                                  const std::string &flow_integrated_filename) const {
    const int nflow = to_order;             // orders 1..to_order
    std::vector<double> normalization(npT_, 0.0);
    std::vector<double> vn(static_cast<size_t>(npT_)*nflow*2, 0.0);    // [pT][order][re, im]
    FILE *f1 = fopen(flow_differential_filename.c_str(), "a");
    FILE *f2 = fopen(flow_integrated_filename.c_str(), "a");
    if (!f1 || !f2) {
        iss_host::error("can not open the flow output files in " + path_);
        exit(-1);
    }
    // differential flow: phi integration per pT (emissionfunction.cpp:913-947)
    for (int i = 0; i < npT_; i++) {
        const double pT = pT_[i];
        const double mT = sqrt(mass*mass + pT*pT);
        for (int j = 0; j < nphi_; j++) {
            const double phi = phi_[j], phi_weight = phi_w_[j];
            const double d = dN[static_cast<size_t>(i)*nphi_ + j];
            normalization[i] += d*phi_weight;
            for (int order = 1; order <= to_order; order++) {
                vn[(static_cast<size_t>(i)*nflow + order - 1)*2] += d*phi_weight*cos(order*phi);
                vn[(static_cast<size_t>(i)*nflow + order - 1)*2 + 1] += d*phi_weight*sin(order*phi);
            }
        }
        normalization[i] = normalization[i] + 1e-30;
        // line: pT, mT - m, dN/(2 pi pT dpT), then (real, imag, norm) per order
        fprintf(f1, "%15.8e  %15.8e  %15.8e  ", pT, mT - mass, normalization[i]/(2.0*M_PI));
        for (int t = 0; t < nflow; t++) {
            const double re = vn[(static_cast<size_t>(i)*nflow + t)*2];
            const double im = vn[(static_cast<size_t>(i)*nflow + t)*2 + 1];
            fprintf(f1, "%15.8e  %15.8e  %15.8e  ", re/normalization[i], im/normalization[i],
                    sqrt(re*re + im*im)/normalization[i]);
        }
        fputc('\n', f1);
    }
    // integrated flow (emissionfunction.cpp:955-975)
    double normalizationi = 0;
    std::vector<double> vni(static_cast<size_t>(nflow)*2, 0.0);
    for (int i = 0; i < npT_; i++) {
        const double pT = pT_[i], pT_weight = pT_w_[i];
        normalizationi += normalization[i]*pT*pT_weight;
        for (int t = 0; t < nflow; t++) {
            vni[2*t] += vn[(static_cast<size_t>(i)*nflow + t)*2]*pT*pT_weight;
            vni[2*t + 1] += vn[(static_cast<size_t>(i)*nflow + t)*2 + 1]*pT*pT_weight;
        }
    }
    // line: order, numerator real, imag, flow real, imag, norm
    fprintf(f2, "%15.8e  %15.8e  %15.8e  %15.8e  %15.8e  %15.8e  \n", 0.0, normalizationi, 0.0, 1.0,
            0.0, 1.0);
    for (int t = 0; t < nflow; t++)
        fprintf(f2, "%15.8e  %15.8e  %15.8e  %15.8e  %15.8e  %15.8e  \n", static_cast<double>(1 + t),
                vni[2*t], vni[2*t + 1], vni[2*t]/normalizationi, vni[2*t + 1]/normalizationi,
                sqrt(vni[2*t]*vni[2*t] + vni[2*t + 1]*vni[2*t + 1])/normalizationi);
    fclose(f1);
    fclose(f2);
}

void GpuSpectra::calculate_dN_dphi_(const std::vector<double> &dN, int monval) const {
    std::vector<double> dN_dphi(nphi_, 0.0);
    for (int i = 0; i < npT_; i++)
        for (int j = 0; j < nphi_; j++)
            dN_dphi[j] += dN[static_cast<size_t>(i)*nphi_ + j]*pT_[i]*pT_w_[i];
    const std::string fn = path_ + "/dN_dphi_" + std::to_string(monval) + ".dat";
    FILE *f = fopen(fn.c_str(), "w");
    if (!f) {
        iss_host::error("can not open " + fn);
        exit(-1);
    }
    // formatedPrint (arsenal.cpp:948-957): "  " + scientific, 10 digits
    for (int j = 0; j < nphi_; j++)
        fprintf(f, "  %.10e  %.10e  %.10e\n", phi_[j], dN_dphi[j], phi_w_[j]);
    fclose(f);
}

// One file of dN matrices for the whole pdg list (zeros for species that were not chosen) and
// one pair of flow files per chosen species.
void GpuSpectra::calculate_dN_pTdpTdphidy_and_flows_4all() {
    std::cout << std::endl
              << "****************************************************************" << std::endl
              << "Function calculate_dN_pTdpTdphidy_and_flows_4all started... " << std::endl;
    const int calculate_dN_dphi = static_cast<int>(paraRdr_->getVal("calculate_dN_dphi"));
    const int to_order = static_cast<int>(paraRdr_->getVal("calculate_vn_to_order"));
    const std::vector<int> &tab = chosen_particles_sampling_table_;
    // species m reuses the table of m - 1 when the two are "the same" (chains included)
    std::vector<int> source(tab.size()), todo;
    for (size_t m = 0; m < tab.size(); m++) {
        if (m > 0 && particles_are_the_same(tab[m], tab[m - 1])) {
            source[m] = source[m - 1];
        } else {
            source[m] = tab[m];
            todo.push_back(tab[m]);
        }
    }
    compute_tables_(todo);
    for (size_t m = 0; m < tab.size(); m++) {
        const particle_info &p = particles_[tab[m]];
        std::cout << "Index: " << m << ", Name: " << p.name << ", Monte-carlo index: " << p.monval
                  << std::endl;
        if (source[m] != tab[m]) {
            std::cout << " -- Using dN_pTdpTdphidy from previous calculation... " << std::endl;
            dN_[tab[m]] = dN_[source[m]];
        } else {
            std::cout << " -- Calculating dN_pTdpTdphidy... " << std::endl;
        }
        if (calculate_dN_dphi) calculate_dN_dphi_(dN_[tab[m]], p.monval);
        const std::string fd = path_ + "/thermal_" + std::to_string(p.monval) + "_vndata.dat";
        const std::string fi = path_ + "/thermal_" + std::to_string(p.monval) + "_integrated__vndata.dat";
        remove(fd.c_str());
        remove(fi.c_str());
        calculate_flows_(dN_[tab[m]], p.mass, to_order, fd, fi);
    }
    const std::string fn = path_ + "/dN_pTdpTdphidy.dat";
    remove(fn.c_str());
    FILE *f = fopen(fn.c_str(), "w");
    if (!f) {
        iss_host::error("can not open " + fn);
        exit(-1);
    }
    for (size_t n = 0; n < particles_.size(); n++) print_dN_table_(f, dN_[n].empty() ? nullptr : &dN_[n]);
    fclose(f);
    info(" -- Calculate_dN_pTdpTdphidy_and_flows_4all finishes.");
}

// Historic (Azspectra) layout: every species of the pdg list in table order, all matrices in one
// file, all flows in v2data.dat / v2data-inte.dat with comment headers.  The working table is one
// buffer: a species that was not chosen zeroes it, a species "the same" as its predecessor in the
// pdg list keeps whatever the buffer holds (emissionfunction.cpp:1062-1087).
void GpuSpectra::calculate_dN_pTdpTdphidy_and_flows_4all_old_output() {
    std::cout << std::endl
              << "*****************************************************************" << std::endl
              << "Function calculate_dN_pTdpTdphidy_and_flows_4all(old) started... " << std::endl;
    const std::string fn = path_ + "/dN_pTdpTdphidy.dat";
    const std::string fd = path_ + "/v2data.dat", fi = path_ + "/v2data-inte.dat";
    remove(fn.c_str());
    remove(fd.c_str());
    remove(fi.c_str());
    const int calculate_dN_dphi = static_cast<int>(paraRdr_->getVal("calculate_dN_dphi"));
    const int to_order = static_cast<int>(paraRdr_->getVal("calculate_vn_to_order"));
    const int Nparticles = static_cast<int>(particles_.size());
    std::vector<int> todo;
    for (int n = 0; n < Nparticles; n++)
        if (chosen_particles_01_table_[n] == 1 && !(n > 0 && particles_are_the_same(n, n - 1)))
            todo.push_back(n);
    compute_tables_(todo);
    const size_t npt = static_cast<size_t>(npT_)*nphi_;
    std::vector<double> buffer(npt, 0.0);
    for (int n = 0; n < Nparticles; n++) {
        const particle_info &p = particles_[n];
        std::cout << "Index: " << n << ", Name: " << p.name << ", Monte-carlo index: " << p.monval;
        if (chosen_particles_01_table_[n] == 0) {
            std::cout << " ...skipped." << std::endl;
            std::fill(buffer.begin(), buffer.end(), 0.0);
        } else {
            std::cout << std::endl;
            if (n > 0 && particles_are_the_same(n, n - 1)) {
                std::cout << " -- Using previously calculated dN_pTdpTdphidy... " << std::endl;
            } else {
                std::cout << " -- Calculating dN_pTdpTdphidy... " << std::endl;
                buffer = dN_[n];
            }
            if (calculate_dN_dphi) calculate_dN_dphi_(buffer, p.monval);
        }
        dN_[n] = buffer;
        FILE *f = fopen(fn.c_str(), "a");
        if (!f) {
            iss_host::error("can not open " + fn);
            exit(-1);
        }
        print_dN_table_(f, &buffer);
        fclose(f);
        FILE *f1 = fopen(fd.c_str(), "a");
        FILE *f2 = fopen(fi.c_str(), "a");
        if (!f1 || !f2) {
            iss_host::error("can not open the flow output files in " + path_);
            exit(-1);
        }
        fprintf(f1, "# Output for particle: %s\n#                 %d\n", p.name.c_str(), p.monval);
        fprintf(f2, "# For: %s\n", p.name.c_str());
        fclose(f1);
        fclose(f2);
        calculate_flows_(buffer, p.mass, to_order, fd, fi);
    }
    info(" -- Calculate_dN_pTdpTdphidy_and_flows_4all finishes.");
}

void GpuSpectra::shell() {
    const int calculate_vn = static_cast<int>(paraRdr_->getVal("calculate_vn"));
    const int historic_format = static_cast<int>(paraRdr_->getVal("use_historic_flow_output_format"));
    // MC_sampling = 2: EmissionFunctionArray::shell computes the spectra first when calculate_vn is
    // set, then samples (emissionfunction.cpp:2554-2572); the sampling half is GpuFSSW's legacy mode
    if (MC_sampling_ != 0 && MC_sampling_ != 2) {
        iss_host::error("the legacy EmissionFunctionArray grid samplers (MC_sampling = 1, 3) are out of "
                        "scope of the B200 engine: use MC_sampling = 4 (FSSW) or 2 (conventional) to "
                        "sample, or MC_sampling = 0 with calculate_vn = 1 for the smooth spectra and flows");
        exit(-1);
    }
    if (calculate_vn) {
        if (historic_format) calculate_dN_pTdpTdphidy_and_flows_4all_old_output();
        else calculate_dN_pTdpTdphidy_and_flows_4all();
    }
}
