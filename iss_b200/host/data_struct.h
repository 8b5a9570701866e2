// data_struct.h -- plain data records of the iSS drop-in API.
//
// Source-compatible with the records callers of the reference see through
// "iSS.h" (reference src/data_struct.h:8-84): same type names, same member
// names and types.  iSS_Hadron additionally keeps the reference's 40-byte
// layout because it is what the binary writer and the device kernels emit.
#ifndef ISS_B200_DATA_STRUCT_H_
#define ISS_B200_DATA_STRUCT_H_

#include <array>
#include <string>
#include <vector>

namespace iSS_data {
constexpr double hbarC = 0.197327053;                // GeV fm (reference data_struct.h:9)
using Vec4 = std::array<float, 4>;
using ViscousVec = std::array<double, 10>;
constexpr int NUMBER_OF_LINES_TO_WRITE = 100000;
constexpr int AMOUNT_OF_OUTPUT = 0;
}  // namespace iSS_data

enum class AfterburnerType { UrQMD, SMASH, JAM, PDG_Decay };

// one line "monval Npart BR d1 .. d5" of a pdg-*.dat decay block
struct decay_channel_info {
    int decay_Npart;
    double branching_ratio;
    int decay_part[5];
};

// one species of a pdg-*.dat table
struct particle_info {
    int monval;
    std::string name;
    double mass;
    double width;
    int gspin;
    int baryon;
    int strange;
    int charm;
    int bottom;
    int gisospin;
    int charge;
    int decays;
    int stable;
    std::vector<decay_channel_info *> decay_channels;
    int sign;       // -1 Bose-Einstein, +1 Fermi-Dirac, 0 Boltzmann
};

// freeze-out cell in Milne coordinates as read from the hydro output
struct FO_surf {
    float tau, xpt, ypt, eta;
    float da0, da1, da2, da3;
    float u0, u1, u2, u3;
    float Edec, Tdec, Pdec;
    float Bn, muB, muS, muQ;
    float pi00, pi01, pi02, pi03, pi11, pi12, pi13, pi22, pi23, pi33;
    float bulkPi;
    float qmu0, qmu1, qmu2, qmu3;
    std::vector<float> particle_mu_PCE;
};

// the same cell in its local rest frame, (t, x, y, z) components
struct FO_surf_LRF {
    float tau, xpt, ypt, eta;
    iSS_data::Vec4 da_mu_LRF;
    iSS_data::Vec4 u_tz;
    float Edec, Tdec, Pdec;
    float Bn, muB, muS, muQ;
    float bulkPi;
    float piLRF_xx, piLRF_xy, piLRF_xz, piLRF_yy, piLRF_yz;
    float qmuLRF_x, qmuLRF_y, qmuLRF_z;
    std::vector<float> particle_mu_PCE;
};

// sampled hadron, 40 bytes
struct iSS_Hadron {
    int pid;
    float mass;
    float E, px, py, pz;
    float t, x, y, z;
};
static_assert(sizeof(iSS_Hadron) == 40, "iSS_Hadron must stay a 40-byte record");

#endif  // ISS_B200_DATA_STRUCT_H_
