// main.cpp -- iSS.e command line of the B200 engine.
// Same contract as the reference CLI (src/main.cpp:22-74):
//   iSS.e [param_file] [work_path] [surface_file] [key=value ...]
// positional arguments are recognised by the absence of '=', every key=value token
// overrides the parameter file, perform_checks=1 writes the QA files after sampling.
#include <iostream>
#include <string>

#include "iSS.h"

int main(int argc, char **argv) {
    std::cout << std::endl
              << "  iSS on B200 -- Cooper-Frye particlization, CUDA sm_100a engine" << std::endl
              << "  (command line and file formats of iSpectraSampler 2.0)" << std::endl
              << std::endl;
    std::string positional[3] = {"iSS_parameters.dat", "results", "surface.dat"};
    for (int i = 1; i <= 3 && i < argc; i++) {
        const std::string tok = argv[i];
        if (tok.find('=') == std::string::npos) positional[i - 1] = tok;
    }
    std::cout << "input file : " << positional[0] << std::endl;
    std::cout << "work folder path : " << positional[1] << std::endl;
    std::cout << "surface filename : " << positional[2] << std::endl;

    iSS sampler(positional[1], "iSS_tables", "iSS_tables", positional[0], positional[2]);
    sampler.paraRdr_ptr->readFromArguments(argc, argv);
    sampler.paraRdr_ptr->echo();

    const int status = sampler.shell();
    if (static_cast<int>(sampler.paraRdr_ptr->getVal("perform_checks", 0)) == 1)
        sampler.perform_checks();
    if (status == 0) std::cout << "Program executed normally." << std::endl;
    return status;
}
