// host_dump.cpp -- runs the HOST half of the pipeline only (no GPU needed): parse music_input,
// surface, HRG regulation, particle table, local-rest-frame transform, species ordering, and
// dumps the result in the format of oracle/ref_driver.cpp so that tests can compare the ingest
// bit for bit with the reference.
//   iss_host_dump <param_file> <work_path> <surface_file> <out_prefix> [key=value ...]
//     -> <out_prefix>.lrf.bin      int64 ncell, ncell x 28 float32 (field order of ISS_F_*)
//        <out_prefix>.species.txt  monval mass gspin baryon strange charge sign stable
//   With MC_sampling != 4 (legacy class: lab-frame cells are kept, src/iSS.cpp:105-109) instead:
//        <out_prefix>.lab.bin      int64 ncell, ncell x 32 float32 (ISS_L_* order of include/iss_cuda.h)
//        <out_prefix>.pos.bin      ncell x 4 float32: xpt, ypt, eta, 0
//        <out_prefix>.species.txt  in the legacy sampling order (sorted by mass only if
//                                  grouping_particles is set, emissionfunction.cpp:190-206)
#include <cstdint>
#include <cstdio>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <string>

#include "../gpu_fssw.h"
#include "../iSS.h"

int main(int argc, char **argv) {
    if (argc < 5) {
        std::cerr << "usage: iss_host_dump param path surface out_prefix [key=value ...]\n";
        return 2;
    }
    const std::string param = argv[1], path = argv[2], surface = argv[3], out = argv[4];
    iSS sampler(path, "iSS_tables", "iSS_tables", param, surface);
    for (int i = 5; i < argc; i++) sampler.paraRdr_ptr->phraseOneLine(argv[i]);
    sampler.read_in_FO_surface();

    const bool legacy = sampler.paraRdr_ptr->getVal("MC_sampling") != 4;
    if (legacy) {
        const auto &lab = sampler.get_lab_surface();
        FILE *fl = fopen((out + ".lab.bin").c_str(), "wb");
        FILE *fp = fopen((out + ".pos.bin").c_str(), "wb");
        if (!fl || !fp) return 1;
        const int64_t nl = static_cast<int64_t>(lab.size());
        fwrite(&nl, sizeof(nl), 1, fl);
        for (const auto &c : lab) {
            const float rec[ISS_LAB_NFIELD] = {
                c.tau, c.u0, c.u1, c.u2, c.u3, c.da0, c.da1, c.da2, c.da3, c.Tdec, c.Pdec, c.Edec,
                c.muB, c.muS, c.muQ, c.pi00, c.pi01, c.pi02, c.pi03, c.pi11, c.pi12, c.pi13, c.pi22,
                c.pi23, c.pi33, c.bulkPi, c.Bn, c.qmu0, c.qmu1, c.qmu2, c.qmu3, 0.f};
            fwrite(rec, sizeof(float), ISS_LAB_NFIELD, fl);
            const float pos[4] = {c.xpt, c.ypt, c.eta, 0.f};
            fwrite(pos, sizeof(float), 4, fp);
        }
        fclose(fl);
        fclose(fp);
        const auto &particles = sampler.get_particle_table();
        const std::vector<int> order = GpuFSSW::order_species(
            sampler.read_chosen_particles(), particles,
            sampler.paraRdr_ptr->getVal("grouping_particles") != 0);
        std::ofstream sp(out + ".species.txt");
        sp << std::setprecision(17);
        for (int idx : order) {
            const particle_info &p = particles[idx];
            sp << p.monval << " " << p.mass << " " << p.gspin << " " << p.baryon << " " << p.strange
               << " " << p.charge << " " << p.sign << " " << p.stable << "\n";
        }
        return 0;
    }

    const auto &surf = sampler.get_LRF_surface();
    FILE *f = fopen((out + ".lrf.bin").c_str(), "wb");
    if (!f) return 1;
    const int64_t n = static_cast<int64_t>(surf.size());
    fwrite(&n, sizeof(n), 1, f);
    for (const auto &s : surf) {
        const float rec[ISS_NFIELD] = {
            s.tau, s.xpt, s.ypt, s.eta,
            s.da_mu_LRF[0], s.da_mu_LRF[1], s.da_mu_LRF[2], s.da_mu_LRF[3],
            s.u_tz[0], s.u_tz[1], s.u_tz[2], s.u_tz[3],
            s.Edec, s.Tdec, s.Pdec, s.Bn, s.muB, s.muS, s.muQ, s.bulkPi,
            s.piLRF_xx, s.piLRF_xy, s.piLRF_xz, s.piLRF_yy, s.piLRF_yz,
            s.qmuLRF_x, s.qmuLRF_y, s.qmuLRF_z};
        fwrite(rec, sizeof(float), ISS_NFIELD, f);
    }
    fclose(f);

    const auto &particles = sampler.get_particle_table();
    const std::vector<int> order =
        GpuFSSW::order_species(sampler.read_chosen_particles(), particles);
    std::ofstream sp(out + ".species.txt");
    sp << std::setprecision(17);
    for (int idx : order) {
        const particle_info &p = particles[idx];
        sp << p.monval << " " << p.mass << " " << p.gspin << " " << p.baryon << " " << p.strange
           << " " << p.charge << " " << p.sign << " " << p.stable << "\n";
    }
    return 0;
}
