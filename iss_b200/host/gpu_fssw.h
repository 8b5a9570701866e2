// gpu_fssw.h -- host driver of the device sampler; takes the place of the reference's
// `class FSSW` (src/FSSW.{h,cpp}) behind `class iSS`.  It owns the CUDA handle
// (include/iss_cuda.h), loads the delta-f coefficient tables with the reference's file
// formats, uploads surface/species/tables, runs yields -> multiplicities -> sampling ->
// decays in event batches sized for the GPU's memory, keeps the hadron lists in one pinned
// host buffer and writes the reference's output formats.
#ifndef ISS_B200_GPU_FSSW_H_
#define ISS_B200_GPU_FSSW_H_

#include <cstdint>
#include <cstdio>
#include <memory>
#include <string>
#include <vector>

#include "../../include/iss_cuda.h"
#include "ParameterReader.h"
#include "data_struct.h"

class GpuFSSW {
 public:
    GpuFSSW(long seed, const std::vector<int> &chosen_monvals,
            const std::vector<particle_info> &particles,
            const std::vector<FO_surf_LRF> &FOsurf_LRF, int flag_PCE, ParameterReader *paraRdr,
            std::string path, std::string table_path, AfterburnerType afterburner_type,
            const float *packed_lrf = nullptr);
    // Legacy mode: the reference's EmissionFunctionArray "conventional" sampler, MC_sampling = 2
    // (src/emissionfunction.cpp:3273-3623), over the LAB-frame cells that iSS::read_in_FO_surface
    // keeps when MC_sampling != 4 (src/iSS.cpp:105-109).  Everything downstream of the sampling
    // kernel (event batches, decays, QA, hadron lists, writers) is shared with the FSSW mode.
    GpuFSSW(long seed, const std::vector<int> &chosen_monvals,
            const std::vector<particle_info> &particles, const std::vector<FO_surf> &FOsurf_lab,
            int flag_PCE, ParameterReader *paraRdr, std::string path, std::string table_path,
            AfterburnerType afterburner_type);
    ~GpuFSSW();

    // chosen list -> indices into the pdg table in sampling order: unknown ids dropped with a
    // warning, stable ascending sort by mass (FSSW.cpp:115-162).  Needs no GPU.
    // FO_surf_LRF records -> [n][ISS_NFIELD] floats in ISS_F_* order (the upload layout), on
    // several threads.  `class iSS` keeps such a block in pinned memory per surface, so that
    // repeated generate_samples() calls copy it to the device without repacking.
    static void pack_surface(const std::vector<FO_surf_LRF> &surf, float *dst, int64_t c0, int64_t c1);
    static std::vector<int> order_species(const std::vector<int> &chosen_monvals,
                                          const std::vector<particle_info> &particles,
                                          bool sort_by_mass = true);

    void shell();       // it all starts here, as in FSSW::shell (FSSW.cpp:344-361)

    // pieces of shell(), public so that hosts/tests can drive them separately
    void compute_yields();
    void sample_events();
    void combine_samples_to_OSCAR();
    void combine_samples_to_gzip_file();
    void combine_samples_to_binary_file();
    void computeAvgTotalEnergyMomentum();

    int get_number_of_sampled_events() const { return static_cast<int>(nev_); }
    int get_number_of_particles(int iev) const {
        return static_cast<int>(event_off_[iev + 1] - event_off_[iev]);
    }
    iSS_Hadron get_hadron(int iev, int ipart) const { return hadrons_[event_off_[iev] + ipart]; }
    std::vector<iSS_Hadron> *get_hadron_list_iev(const int iev);

    // B200-engine additions
    iss_handle *cuda_handle() { return h_; }
    const std::vector<iss_species> &species() const { return species_; }
    const std::vector<double> &species_dN() const { return dN_species_; }
    const std::vector<double> &qa_block() const { return qa_; }
    // events behind the QA block: this process's, or those of all ranks when the block was reduced
    // over the ranks of the job (parameter reduce_checks_over_ranks = 1)
    double qa_events() const { return qa_.empty() ? 0. : qa_[0]; }
    int qa_ranks() const { return qa_ranks_; }
    const iSS_Hadron *hadron_buffer() const { return hadrons_; }
    const std::vector<int64_t> &event_offsets() const { return event_off_; }
    int number_of_chosen_particles() const { return static_cast<int>(species_.size()); }

 private:
    ParameterReader *paraRdr_;
    const std::string path_, table_path_;
    const AfterburnerType afterburner_type_;
    const std::vector<particle_info> &particles_;
    const std::vector<FO_surf_LRF> &surf_;
    const float *packed_lrf_ = nullptr;     // optional: surf_ already packed in pinned memory
    const std::vector<FO_surf> *lab_surf_ = nullptr;    // legacy mode (MC_sampling = 2)
    bool legacy_ = false;
    long seed_;
    int hydro_mode_;
    int include_shear_, include_bulk_, include_diff_, bulk_kind_;
    int number_of_repeated_sampling_;
    int flag_perform_decays_, flag_spectators_;
    int use_oscar_, use_gzip_, use_binary_;

    iss_handle *h_ = nullptr;
    int device_ = 0;
    std::vector<iss_species> species_;
    std::vector<int> species_table_idx_;    // index into particles_ (chosen_particles_sampling_table)
    std::vector<double> dN_species_;
    std::vector<double> qa_;
    int qa_ranks_ = 1;
    void add_spectators_to_qa_(const int32_t *pids, int npid);
    // per-species text files of the legacy class (output_samples_into_files = 1)
    int flag_sample_files_ = 0;
    std::vector<FILE *> sample_files_, control_files_;
    void begin_sample_files_();
    void append_sample_files_(int64_t nev_batch, int64_t n_hadrons);
    void end_sample_files_();
    void join_ranks_();     // NCCL communicator of the job on the pooled handle (once per process)

    iSS_Hadron *hadrons_ = nullptr;         // pinned, all events
    int64_t hadron_cap_ = 0;
    int64_t nev_ = 0;
    std::vector<int64_t> event_off_;
    std::vector<std::unique_ptr<std::vector<iSS_Hadron>>> event_cache_;
    std::vector<iSS_Hadron> spectators_;

    void init_(const std::vector<int> &chosen_monvals, int flag_PCE);
    void upload_lab_surface_();
    void check_(int rc, const char *what);
    void select_species_(const std::vector<int> &chosen_monvals);
    void upload_surface_();
    void upload_tables_();
    void upload_decay_table_();
    void read_spectators_(const std::string &file);
    void reserve_hadrons_(int64_t need);
    int compute_number_of_sampling_needed_(long number_of_particles_needed);
};

#endif  // ISS_B200_GPU_FSSW_H_
