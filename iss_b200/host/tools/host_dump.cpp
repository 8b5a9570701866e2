// host_dump.cpp -- runs the HOST half of the pipeline only (no GPU needed): parse music_input,
// surface, HRG regulation, particle table, local-rest-frame transform, species ordering, and
// dumps the result in the format of oracle/ref_driver.cpp so that tests can compare the ingest
// bit for bit with the reference.
//   iss_host_dump <param_file> <work_path> <surface_file> <out_prefix> [key=value ...]
//     -> <out_prefix>.lrf.bin      int64 ncell, ncell x 28 float32 (field order of ISS_F_*)
//        <out_prefix>.species.txt  monval mass gspin baryon strange charge sign stable
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
