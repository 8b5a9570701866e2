// oracle/ref_driver.cpp -- TEST INFRASTRUCTURE ONLY.
//
// Drives the UNMODIFIED reference sources (compiled where they lie under
// /root/reference/src by oracle/Makefile, objects only into oracle/_ref/) and
// dumps intermediate results of the hot path so that the CUDA product and the
// numpy restatement (oracle/iss_oracle.py) can be pinned against the reference
// itself.  Nothing in the product links or executes this file.
//
//   ref_driver yields  <param_file> <work_path> <surface_file> <out_prefix> [key=value ...]
//       -> <out_prefix>.lrf.bin      int64 ncell, then ncell x 28 float32 (FO_surf_LRF order below)
//          <out_prefix>.species.txt  one line per chosen species in sampling order:
//                                    monval mass gspin baryon strange charge sign stable
//          <out_prefix>.yields.bin   int64 nspecies, int64 ncell, then nspecies x ncell float64
//                                    = FSSW::dN_dxtdy_for_one_particle_species (FSSW.cpp:565-715)
//   ref_driver momentum <m> <T> <mu> <sign> <n> <seed> <out.bin>
//       -> n float64 |p| samples of MomentumSamplerShell::Sample_a_momentum (MomentumSamplerShell.cpp:24)
//   ref_driver decay <table_path> <afterburner 1|2> <pid> <n> <seed> <out.bin>
//       -> n mothers at rest-ish (fixed momentum) decayed once with particle_decay::perform_decays
//          (particle_decay.cpp:265); records: int32 ndaughters then ndaughters x iSS_Hadron (40 B)
#include <string>
#include <vector>
#include <memory>
#include <array>
#include <sstream>
#include <iostream>
#include <fstream>
#include <random>
#include <cmath>
#include <iomanip>
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <gsl/gsl_rng.h>
#include <gsl/gsl_randist.h>

#define private public
#include "iSS.h"
#undef private

static void write_lrf(const std::vector<FO_surf_LRF> &surf, const std::string &fn) {
    FILE *f = fopen(fn.c_str(), "wb");
    int64_t n = surf.size();
    fwrite(&n, sizeof(n), 1, f);
    for (auto const &s : surf) {
        float rec[28] = {
            s.tau, s.xpt, s.ypt, s.eta,
            s.da_mu_LRF[0], s.da_mu_LRF[1], s.da_mu_LRF[2], s.da_mu_LRF[3],
            s.u_tz[0], s.u_tz[1], s.u_tz[2], s.u_tz[3],
            s.Edec, s.Tdec, s.Pdec, s.Bn, s.muB, s.muS, s.muQ, s.bulkPi,
            s.piLRF_xx, s.piLRF_xy, s.piLRF_xz, s.piLRF_yy, s.piLRF_yz,
            s.qmuLRF_x, s.qmuLRF_y, s.qmuLRF_z};
        fwrite(rec, sizeof(float), 28, f);
    }
    fclose(f);
}

static int run_yields(int argc, char **argv) {
    if (argc < 6) { std::cerr << "usage: yields param path surface out_prefix [k=v]\n"; return 2; }
    std::string param = argv[2], path = argv[3], surface = argv[4], out = argv[5];
    iSS sampler(path, "iSS_tables", "iSS_tables", param, surface);
    for (int i = 6; i < argc; i++) sampler.paraRdr_ptr->phraseOneLine(argv[i]);
    sampler.read_in_FO_surface();
    sampler.set_random_seed(1);
    write_lrf(sampler.FOsurf_LRF_array_, out + ".lrf.bin");

    // same construction as iSS::generate_samples (iSS.cpp:130-149)
    Table chosen_particles;
    if (sampler.afterburner_type_ == AfterburnerType::SMASH) {
        chosen_particles.loadTableFromFile("iSS_tables/chosen_particles_SMASH.dat");
    } else if (sampler.afterburner_type_ == AfterburnerType::UrQMD) {
        chosen_particles.loadTableFromFile("iSS_tables/chosen_particles_urqmd_v3.3+.dat");
    } else {
        chosen_particles.loadTableFromFile("iSS_tables/chosen_particles_s95p-v1.dat");
    }
    FSSW fssw(sampler.ran_gen_ptr_, &chosen_particles, sampler.particle_,
              sampler.FOsurf_LRF_array_, sampler.flag_PCE_, sampler.paraRdr_ptr,
              path, "iSS_tables", sampler.afterburner_type_);

    int fix_c0 = static_cast<int>(sampler.paraRdr_ptr->getVal("oracle_fix_14mom_c0", 0));
    if (fix_c0 == 1 && fssw.INCLUDE_BULK_DELTAF == 1 && fssw.bulk_deltaf_kind_ == 11) {
        // FSSW::load_bulk_deltaf_14mom_table (FSSW.cpp:1246-1256) leaves the
        // c0 table uninitialised; overwrite it with the correctly parsed file so that
        // kind 11 has a well-defined oracle (documented deviation).
        std::string folder = (sampler.afterburner_type_ == AfterburnerType::SMASH)
                                 ? "/smash_box" : "/urqmd";
        std::ifstream c0("iSS_tables/deltaf_tables" + folder + "/c0.dat");
        std::string dummy;
        for (int i = 0; i < 3; i++) std::getline(c0, dummy);
        double t1, t2;
        for (int j = 0; j < fssw.deltaf_bulk_coeff_14mom_table_length_mu_; j++)
            for (int i = 0; i < fssw.deltaf_bulk_coeff_14mom_table_length_T_; i++)
                c0 >> t1 >> t2 >> fssw.deltaf_bulk_coeff_14mom_c0_tb_[i][j];
    }

    int ns = fssw.number_of_chosen_particles;
    int64_t ncell = sampler.FOsurf_LRF_array_.size();
    std::ofstream sp(out + ".species.txt");
    sp << std::setprecision(17);
    FILE *f = fopen((out + ".yields.bin").c_str(), "wb");
    int64_t ns64 = ns;
    fwrite(&ns64, sizeof(ns64), 1, f);
    fwrite(&ncell, sizeof(ncell), 1, f);
    for (int n = 0; n < ns; n++) {
        int idx = fssw.chosen_particles_sampling_table[n];
        const particle_info &p = fssw.particles[idx];
        sp << p.monval << " " << p.mass << " " << p.gspin << " " << p.baryon << " "
           << p.strange << " " << p.charge << " " << p.sign << " " << p.stable << "\n";
        fssw.calculate_dN_dxtdy_for_one_particle_species(idx);
        fwrite(fssw.dN_dxtdy_for_one_particle_species.data(), sizeof(double), ncell, f);
    }
    fclose(f);
    return 0;
}

static int run_momentum(int argc, char **argv) {
    if (argc < 9) { std::cerr << "usage: momentum m T mu sign n seed out\n"; return 2; }
    double m = atof(argv[2]), T = atof(argv[3]), mu = atof(argv[4]);
    int sign = atoi(argv[5]);
    long n = atol(argv[6]);
    long seed = atol(argv[7]);
    auto ran = std::make_shared<RandomUtil::Random>(seed);
    MomentumSamplerShell shell(ran);
    std::vector<double> out(n);
    for (long i = 0; i < n; i++) out[i] = shell.Sample_a_momentum(m, T, mu, sign);
    FILE *f = fopen(argv[8], "wb");
    fwrite(out.data(), sizeof(double), n, f);
    fclose(f);
    return 0;
}

static int run_decay(int argc, char **argv) {
    if (argc < 8) { std::cerr << "usage: decay table_path afterburner pid n seed out\n"; return 2; }
    std::string table_path = argv[2];
    AfterburnerType ab = (atoi(argv[3]) == 2) ? AfterburnerType::SMASH : AfterburnerType::UrQMD;
    int pid = atoi(argv[4]);
    long n = atol(argv[5]);
    long seed = atol(argv[6]);
    auto ran = std::make_shared<RandomUtil::Random>(seed);
    particle_decay decayer(ran, ab, table_path);
    FILE *f = fopen(argv[7], "wb");
    for (long i = 0; i < n; i++) {
        iSS_Hadron mother;
        mother.pid = pid;
        mother.mass = decayer.get_particle_mass(pid);
        mother.px = 0.3f; mother.py = -0.2f; mother.pz = 0.5f;
        mother.E = std::sqrt(mother.mass*mother.mass + mother.px*mother.px
                             + mother.py*mother.py + mother.pz*mother.pz);
        mother.t = 1.f; mother.x = 0.5f; mother.y = -0.5f; mother.z = 0.25f;
        std::vector<iSS_Hadron> daughters;
        decayer.perform_decays(&mother, &daughters);
        int32_t nd = daughters.size();
        fwrite(&nd, sizeof(nd), 1, f);
        fwrite(daughters.data(), sizeof(iSS_Hadron), nd, f);
    }
    fclose(f);
    return 0;
}

// ref_driver writers <param_file> <work_path> <surface_file> <hadrons.bin>
//   hadrons.bin: int64 nev, int64 offsets[nev+1], then iSS_Hadron records (40 B).  The list is put
//   into FSSW::Hadron_list and the reference's three writers are called; OSCAR.DAT,
//   particle_samples.gz and particle_samples.bin appear in the current directory.
static int run_writers(int argc, char **argv) {
    if (argc < 6) { std::cerr << "usage: writers param path surface hadrons.bin\n"; return 2; }
    std::string param = argv[2], path = argv[3], surface = argv[4];
    iSS sampler(path, "iSS_tables", "iSS_tables", param, surface);
    for (int i = 6; i < argc; i++) sampler.paraRdr_ptr->phraseOneLine(argv[i]);
    sampler.read_in_FO_surface();
    sampler.set_random_seed(1);
    Table chosen_particles;
    chosen_particles.loadTableFromFile("iSS_tables/chosen_particles_urqmd_v3.3+.dat");
    FSSW fssw(sampler.ran_gen_ptr_, &chosen_particles, sampler.particle_,
              sampler.FOsurf_LRF_array_, sampler.flag_PCE_, sampler.paraRdr_ptr,
              path, "iSS_tables", sampler.afterburner_type_);
    FILE *f = fopen(argv[5], "rb");
    int64_t nev = 0;
    if (fread(&nev, sizeof(nev), 1, f) != 1) return 1;
    std::vector<int64_t> off(nev + 1);
    if (fread(off.data(), sizeof(int64_t), nev + 1, f) != static_cast<size_t>(nev + 1)) return 1;
    for (int64_t ev = 0; ev < nev; ev++) {
        auto *v = new std::vector<iSS_Hadron>(off[ev + 1] - off[ev]);
        if (!v->empty() && fread(v->data(), sizeof(iSS_Hadron), v->size(), f) != v->size()) return 1;
        fssw.Hadron_list->push_back(v);
    }
    fclose(f);
    fssw.combine_samples_to_OSCAR();
    fssw.combine_samples_to_gzip_file();
    fssw.combine_samples_to_binary_file();
    return 0;
}

// ref_driver spectra <param_file> <work_path> <surface_file> <out_prefix> species=<monval,...> [key=value ...]
//   MC_sampling is forced to 0 so that iSS::read_in_FO_surface keeps the lab-frame (Milne) cells
//   (iSS.cpp:105-109).  For every requested species EmissionFunctionArray::calculate_dN_pTdpTdphidy
//   (emissionfunction.cpp:624-829) and calculate_flows (:875-1016) are called.
//   -> <out_prefix>.lab.bin    int64 ncell, then ncell x 32 float32 in the ISS_L_* order of
//                              include/iss_cuda.h (tau, u0-3, da0-3, T, P, e, muB, muS, muQ,
//                              pi00 01 02 03 11 12 13 22 23 33, bulkPi, Bn, q0-3, 0)
//      <out_prefix>.dN.bin     int64 ns, npT, nphi, then per species [npT][nphi] float64 dN and the
//                              same for dN_max
//      <out_prefix>.species.txt  monval mass gspin baryon strange charge sign
//      <out_prefix>.vndiff.<monval>.dat / .vninte.<monval>.dat  the reference's flow files
static int run_spectra(int argc, char **argv) {
    if (argc < 7) { std::cerr << "usage: spectra param path surface out_prefix species=.. [k=v]\n"; return 2; }
    std::string param = argv[2], path = argv[3], surface = argv[4], out = argv[5];
    iSS sampler(path, "iSS_tables", "iSS_tables", param, surface);
    std::vector<int> wanted;
    for (int i = 6; i < argc; i++) {
        std::string a = argv[i];
        if (a.rfind("species=", 0) == 0) {
            std::stringstream ss(a.substr(8));
            std::string tok;
            while (std::getline(ss, tok, ',')) wanted.push_back(atoi(tok.c_str()));
        } else {
            sampler.paraRdr_ptr->phraseOneLine(argv[i]);
        }
    }
    sampler.paraRdr_ptr->setVal("MC_sampling", 0);
    sampler.read_in_FO_surface();
    sampler.set_random_seed(1);
    {
        FILE *f = fopen((out + ".lab.bin").c_str(), "wb");
        int64_t n = sampler.FOsurf_array_.size();
        fwrite(&n, sizeof(n), 1, f);
        for (auto const &c : sampler.FOsurf_array_) {
            float rec[32] = {c.tau, c.u0, c.u1, c.u2, c.u3, c.da0, c.da1, c.da2, c.da3,
                             c.Tdec, c.Pdec, c.Edec, c.muB, c.muS, c.muQ,
                             c.pi00, c.pi01, c.pi02, c.pi03, c.pi11, c.pi12, c.pi13, c.pi22, c.pi23,
                             c.pi33, c.bulkPi, c.Bn, c.qmu0, c.qmu1, c.qmu2, c.qmu3, 0.f};
            fwrite(rec, sizeof(float), 32, f);
        }
        fclose(f);
    }
    // same construction as iSS::generate_samples (iSS.cpp:150-163)
    Table chosen_particles;
    if (sampler.afterburner_type_ == AfterburnerType::SMASH) {
        chosen_particles.loadTableFromFile("iSS_tables/chosen_particles_SMASH.dat");
    } else if (sampler.afterburner_type_ == AfterburnerType::UrQMD) {
        chosen_particles.loadTableFromFile("iSS_tables/chosen_particles_urqmd_v3.3+.dat");
    } else {
        chosen_particles.loadTableFromFile("iSS_tables/chosen_particles_s95p-v1.dat");
    }
    Table pT_tab("iSS_tables/bin_tables/pT_gauss_table.dat");
    Table phi_tab("iSS_tables/bin_tables/phi_gauss_table.dat");
    Table eta_tab("iSS_tables/bin_tables/eta_uni_table.dat");
    EmissionFunctionArray efa(sampler.ran_gen_ptr_, &chosen_particles, &pT_tab, &phi_tab, &eta_tab,
                              sampler.particle_, sampler.FOsurf_array_, sampler.flag_PCE_,
                              sampler.paraRdr_ptr, path, "iSS_tables", sampler.afterburner_type_);
    const int to_order = sampler.paraRdr_ptr->getVal("calculate_vn_to_order");
    const int npT = efa.pT_tab_length, nphi = efa.phi_tab_length;
    std::ofstream sp(out + ".species.txt");
    sp << std::setprecision(17);
    FILE *f = fopen((out + ".dN.bin").c_str(), "wb");
    int64_t hdr[3] = {static_cast<int64_t>(wanted.size()), npT, nphi};
    fwrite(hdr, sizeof(int64_t), 3, f);
    for (int monval : wanted) {
        int idx = -1;
        for (int n = 0; n < efa.Nparticles; n++)
            if (efa.particles[n].monval == monval) idx = n;
        if (idx < 0) { std::cerr << "unknown species " << monval << "\n"; return 1; }
        const particle_info &p = efa.particles[idx];
        sp << p.monval << " " << p.mass << " " << p.gspin << " " << p.baryon << " "
           << p.strange << " " << p.charge << " " << p.sign << "\n";
        efa.calculate_dN_pTdpTdphidy(idx);
        std::vector<double> buf(2*npT*nphi);
        for (int i = 0; i < npT; i++)
            for (int j = 0; j < nphi; j++) {
                buf[i*nphi + j] = efa.dN_pTdpTdphidy->get(i + 1, j + 1);
                buf[npT*nphi + i*nphi + j] = efa.dN_pTdpTdphidy_max->get(i + 1, j + 1);
            }
        fwrite(buf.data(), sizeof(double), buf.size(), f);
        std::string fd = out + ".vndiff." + std::to_string(monval) + ".dat";
        std::string fi = out + ".vninte." + std::to_string(monval) + ".dat";
        remove(fd.c_str());
        remove(fi.c_str());
        efa.calculate_flows(to_order, fd, fi);
    }
    fclose(f);
    return 0;
}

// ref_driver legacy <param_file> <work_path> <surface_file> <out_prefix> [key=value ...]
//   MC_sampling is forced to 2 (EmissionFunctionArray "conventional" sampler,
//   emissionfunction.cpp:3273-3623).  Dumps what that path computes before it draws anything:
//   -> <out_prefix>.lab.bin     int64 ncell, ncell x 32 float32 (ISS_L_* order as in `spectra`)
//      <out_prefix>.pos.bin     ncell x 4 float32: xpt, ypt, eta, 0
//      <out_prefix>.species.txt sampling order (chosen_particles_sampling_table):
//                               monval mass gspin baryon strange charge sign
//      <out_prefix>.yields.bin  int64 ns, ncell, then ns x ncell float64 = dN_dxtdy_for_one_particle_species
//                               (calculate_dN_dxtdy_for_one_particle_species, :2977-3097; NOT clamped)
//      <out_prefix>.max.bin     ns x ncell float64 = estimate_maximum (:4309-4421)
static int run_legacy(int argc, char **argv) {
    if (argc < 6) { std::cerr << "usage: legacy param path surface out_prefix [k=v]\n"; return 2; }
    std::string param = argv[2], path = argv[3], surface = argv[4], out = argv[5];
    iSS sampler(path, "iSS_tables", "iSS_tables", param, surface);
    for (int i = 6; i < argc; i++) sampler.paraRdr_ptr->phraseOneLine(argv[i]);
    sampler.paraRdr_ptr->setVal("MC_sampling", 2);
    sampler.read_in_FO_surface();
    sampler.set_random_seed(1);
    {
        FILE *f = fopen((out + ".lab.bin").c_str(), "wb");
        FILE *g = fopen((out + ".pos.bin").c_str(), "wb");
        int64_t n = sampler.FOsurf_array_.size();
        fwrite(&n, sizeof(n), 1, f);
        for (auto const &c : sampler.FOsurf_array_) {
            float rec[32] = {c.tau, c.u0, c.u1, c.u2, c.u3, c.da0, c.da1, c.da2, c.da3,
                             c.Tdec, c.Pdec, c.Edec, c.muB, c.muS, c.muQ,
                             c.pi00, c.pi01, c.pi02, c.pi03, c.pi11, c.pi12, c.pi13, c.pi22, c.pi23,
                             c.pi33, c.bulkPi, c.Bn, c.qmu0, c.qmu1, c.qmu2, c.qmu3, 0.f};
            fwrite(rec, sizeof(float), 32, f);
            float pos[4] = {c.xpt, c.ypt, c.eta, 0.f};
            fwrite(pos, sizeof(float), 4, g);
        }
        fclose(f);
        fclose(g);
    }
    Table chosen_particles;
    if (sampler.afterburner_type_ == AfterburnerType::SMASH) {
        chosen_particles.loadTableFromFile("iSS_tables/chosen_particles_SMASH.dat");
    } else if (sampler.afterburner_type_ == AfterburnerType::UrQMD) {
        chosen_particles.loadTableFromFile("iSS_tables/chosen_particles_urqmd_v3.3+.dat");
    } else {
        chosen_particles.loadTableFromFile("iSS_tables/chosen_particles_s95p-v1.dat");
    }
    Table pT_tab("iSS_tables/bin_tables/pT_gauss_table.dat");
    Table phi_tab("iSS_tables/bin_tables/phi_gauss_table.dat");
    Table eta_tab("iSS_tables/bin_tables/eta_uni_table.dat");
    EmissionFunctionArray efa(sampler.ran_gen_ptr_, &chosen_particles, &pT_tab, &phi_tab, &eta_tab,
                              sampler.particle_, sampler.FOsurf_array_, sampler.flag_PCE_,
                              sampler.paraRdr_ptr, path, "iSS_tables", sampler.afterburner_type_);
    TableFunction z_exp_m_z(std::string("iSS_tables/z_exp_m_z.dat"));
    z_exp_m_z.interpolation_model = 5;
    const int64_t ns = efa.number_of_chosen_particles, nc = efa.FO_length;
    std::ofstream sp(out + ".species.txt");
    sp << std::setprecision(17);
    FILE *fy = fopen((out + ".yields.bin").c_str(), "wb");
    FILE *fm = fopen((out + ".max.bin").c_str(), "wb");
    int64_t hdr[2] = {ns, nc};
    fwrite(hdr, sizeof(int64_t), 2, fy);
    std::vector<double> mx(nc);
    for (int64_t n = 0; n < ns; n++) {
        const int idx = efa.chosen_particles_sampling_table[n];
        const particle_info &p = efa.particles[idx];
        sp << p.monval << " " << p.mass << " " << p.gspin << " " << p.baryon << " "
           << p.strange << " " << p.charge << " " << p.sign << "\n";
        efa.calculate_dN_dxtdy_for_one_particle_species(idx);
        fwrite(efa.dN_dxtdy_for_one_particle_species.data(), sizeof(double), nc, fy);
        for (int64_t l = 0; l < nc; l++) {
            const FO_surf *surf = &sampler.FOsurf_array_[l];
            std::array<double, 3> bulk = {0.0};
            if (efa.INCLUDE_BULK_DELTAF == 1) efa.getbulkvisCoefficients(surf->Tdec, bulk);
            double kq = 1.0;
            if (efa.INCLUDE_DIFFUSION_DELTAF == 1) kq = efa.get_deltaf_qmu_coeff(surf->Tdec, surf->muB);
            mx[l] = efa.estimate_maximum(surf, idx, p.mass, p.sign, p.gspin, p.baryon, p.strange,
                                         p.charge, z_exp_m_z, bulk, kq);
        }
        fwrite(mx.data(), sizeof(double), nc, fm);
    }
    fclose(fy);
    fclose(fm);
    return 0;
}

int main(int argc, char **argv) {
    if (argc < 2) { std::cerr << "usage: ref_driver yields|momentum|decay ...\n"; return 2; }
    std::string mode = argv[1];
    if (mode == "yields") return run_yields(argc, argv);
    if (mode == "momentum") return run_momentum(argc, argv);
    if (mode == "decay") return run_decay(argc, argv);
    if (mode == "writers") return run_writers(argc, argv);
    if (mode == "spectra") return run_spectra(argc, argv);
    if (mode == "legacy") return run_legacy(argc, argv);
    std::cerr << "unknown mode " << mode << "\n";
    return 2;
}
