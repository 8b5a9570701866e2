// gpu_spectra.h -- host driver of the smooth Cooper-Frye spectra; takes the place of the spectra /
// flow half of the reference's legacy `class EmissionFunctionArray` (src/emissionfunction.{h,cpp})
// behind `class iSS` when MC_sampling == 0 and calculate_vn == 1.  The dense sum over
// cells x (y - eta_s) x pT x phi runs on the GPU (iss_cuda_spectra); the flow harmonics, the
// grouping of species with equal quantum numbers and the output files are host work.
// The conventional sampler of that class (MC_sampling = 2) is GpuFSSW's legacy mode (gpu_fssw.h); its
// grid samplers (MC_sampling = 1, 3) are out of scope.
#ifndef ISS_B200_GPU_SPECTRA_H_
#define ISS_B200_GPU_SPECTRA_H_

#include <string>
#include <vector>

#include "../../include/iss_cuda.h"
#include "ParameterReader.h"
#include "data_struct.h"

class GpuSpectra {
 public:
    // the three bin tables are read from <table_path>/bin_tables as iSS::generate_samples does
    // (iSS.cpp:151-156): pT_gauss_table.dat, phi_gauss_table.dat, eta_uni_table.dat
    GpuSpectra(const std::vector<int> &chosen_monvals, const std::vector<particle_info> &particles,
               const std::vector<FO_surf> &FOsurf, int flag_PCE, ParameterReader *paraRdr,
               std::string path, std::string table_path, AfterburnerType afterburner_type);
    ~GpuSpectra();

    void shell();       // EmissionFunctionArray::shell (emissionfunction.cpp:2542-2591)

    // emissionfunction.cpp:1129-1221 and :1036-1120
    void calculate_dN_pTdpTdphidy_and_flows_4all();
    void calculate_dN_pTdpTdphidy_and_flows_4all_old_output();
    // emissionfunction.cpp:2511-2537
    bool particles_are_the_same(int idx1, int idx2) const;

    // B200-engine additions: the tables of the last run, [npT][nphi] per pdg-table index
    // (empty when the species was not calculated)
    const std::vector<double> &dN_pTdpTdphidy(int particle_idx) const { return dN_[particle_idx]; }
    int pT_tab_length() const { return npT_; }
    int phi_tab_length() const { return nphi_; }
    const std::vector<int> &sampling_table() const { return chosen_particles_sampling_table_; }
    double last_kernel_ms() const { return kernel_ms_; }
    double last_evaluations() const { return evaluations_; }

 private:
    ParameterReader *paraRdr_;
    const std::string path_, table_path_;
    const std::vector<particle_info> &particles_;
    const std::vector<FO_surf> &surf_;
    int include_shear_, include_bulk_, include_diff_, bulk_kind_;
    int restrict_deltaf_, use_pos_dN_only_, grouping_particles_;
    double deltaf_max_ratio_, grouping_tolerance_;
    int MC_sampling_;

    std::vector<double> pT_, pT_w_, phi_, phi_w_, eta_, eta_w_;
    int npT_ = 0, nphi_ = 0, neta_ = 0;
    std::vector<int> chosen_particles_01_table_;
    std::vector<int> chosen_particles_sampling_table_;
    std::vector<std::vector<double>> dN_;       // by pdg-table index

    iss_handle *h_ = nullptr;
    int device_ = 0;
    double kernel_ms_ = 0., evaluations_ = 0.;

    void check_(int rc, const char *what);
    void upload_surface_();
    void upload_tables_();
    // GPU: dN tables of the listed pdg-table indices (one batched call)
    void compute_tables_(const std::vector<int> &particle_idx);
    // emissionfunction.cpp:875-1016; appends to the two files
    void calculate_flows_(const std::vector<double> &dN, double mass, int to_order,
                          const std::string &flow_differential_filename,
                          const std::string &flow_integrated_filename) const;
    // emissionfunction.cpp:2428-2455
    void calculate_dN_dphi_(const std::vector<double> &dN, int monval) const;
    // Table::printTable of a [cols = npT][rows = nphi] table (Table.cpp:166-176)
    void print_dN_table_(FILE *f, const std::vector<double> *dN) const;
};

#endif  // ISS_B200_GPU_SPECTRA_H_
