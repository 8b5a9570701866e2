// dropin_host.cpp -- a host written against the REFERENCE's public API only (src/iSS.h:16-102,
// the call sequence of src/main.cpp:58-69 and of MUSIC/JETSCAPE wrappers): it must compile and
// link unchanged against this repository's iSS.h / libiSS.so.
//   dropin_host <work_path> <param_file> <surface_file> <nev> <seed>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "iSS.h"

int main(int argc, char **argv) {
    if (argc < 6) return 2;
    const std::string path = argv[1], param = argv[2], surface = argv[3];
    iSS sampler(path, "iSS_tables", "iSS_tables", param, surface);
    sampler.paraRdr_ptr->setVal("number_of_repeated_sampling", atof(argv[4]));
    sampler.paraRdr_ptr->setVal("perform_decays", 0);
    for (int i = 6; i < argc; i++) sampler.paraRdr_ptr->phraseOneLine(argv[i]);
    if (sampler.read_in_FO_surface() != 0) return 3;
    sampler.set_random_seed(atoi(argv[5]));
    if (sampler.generate_samples() != 0) return 4;
    const int nev = sampler.get_number_of_sampled_events();
    long total = 0;
    double E = 0.;
    for (int iev = 0; iev < nev; iev++) {
        std::vector<iSS_Hadron> *list = sampler.get_hadron_list_iev(iev);
        if (static_cast<int>(list->size()) != sampler.get_number_of_particles(iev)) return 5;
        total += static_cast<long>(list->size());
        for (const iSS_Hadron &h : *list) E += h.E;
        if (!list->empty()) {
            const iSS_Hadron first = sampler.get_hadron(iev, 0);
            if (first.pid != (*list)[0].pid || first.E != (*list)[0].E) return 6;
        }
    }
    printf("DROPIN events %d hadrons %ld mean_E %.6f\n", nev, total, total ? E/total : 0.);
    sampler.clear();
    return 0;
}
