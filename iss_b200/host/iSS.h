// iSS.h -- drop-in facade of the B200 Cooper-Frye particlization engine.
//
// Same public class API as the reference's `class iSS` (reference src/iSS.h:16-102):
// constructor arguments and defaults, the public `paraRdr_ptr`, set_random_seed,
// perform_checks, shell / read_in_FO_surface / generate_samples, the event/hadron
// accessors, clear() and transform_to_local_rest_frame.  Host code only parses input
// and drives the device: yields, multiplicities, momentum sampling, boost, decays and
// QA histograms run in the CUDA library behind include/iss_cuda.h.  There is no CPU
// sampler in this library: MC_sampling must be 4 (FSSW path), 2 (legacy conventional
// sampler) or 0 (smooth spectra and flows) and a B200-class GPU
// must be present, otherwise the call prints a message and exits like the reference's
// other fatal errors (return 0 on success, exit(+-1) on failure; no exceptions).
#ifndef ISS_H
#define ISS_H

#include <array>
#include <memory>
#include <string>
#include <vector>

#include "ParameterReader.h"
#include "data_struct.h"

class GpuFSSW;
class GpuSpectra;

class iSS {
 private:
    const std::string path_;
    const std::string table_path_;
    const std::string particle_table_path_;
    const std::string surface_filename_;

    std::vector<FO_surf> FOsurf_array_;         // lab-frame cells, kept when MC_sampling != 4
    std::vector<FO_surf_LRF> FOsurf_LRF_array_;
    std::vector<float> FOsurf_Tmunu_;
    std::vector<float> FOsurf_Q_;

    int flag_PCE_;
    AfterburnerType afterburner_type_;
    long randomSeed_;
    bool seed_set_;

    std::vector<particle_info> particle_;
    std::unique_ptr<GpuFSSW> spectra_sampler_;
    std::unique_ptr<GpuSpectra> efa_;           // smooth spectra / flows (MC_sampling = 0)

    void require_fssw_() const;
    void require_supported_mode_() const;
    void accumulate_Tmunu_(const std::vector<FO_surf> &cells);
    void report_Tmunu_() const;
    void ingest_binary_on_device_(class read_FOdata &reader, int64_t nbin);
    // FOsurf_LRF_array_ in the upload layout ([n][28] floats) in pinned host memory, built once per
    // surface (engine addition: every generate_samples() copies it to the device as it is)
    void *lrf_packed_ = nullptr;
    int64_t lrf_packed_bytes_ = 0, lrf_packed_n_ = -1;
    void ensure_packed_lrf_();
    void drop_packed_lrf_();

 public:
    iSS(std::string path, std::string table_path = "iSS_tables",
        std::string particle_table_path = "iSS_tables",
        std::string inputfile = "iSS_parameters.dat",
        std::string surface_filename = "surface.dat");
    ~iSS();

    ParameterReader *paraRdr_ptr;

    void set_random_seed();
    void set_random_seed(int randomSeed_in);

    void perform_checks();
    void construct_Tmunu_from_particle_samples();

    int shell();
    int read_in_FO_surface();
    int generate_samples();

    int get_number_of_sampled_events();
    int get_number_of_particles(int iev);
    iSS_Hadron get_hadron(int iev, int ipart);
    std::vector<iSS_Hadron> *get_hadron_list_iev(const int iev);

    void clear();

    // Milne-frame cells -> local rest frame in (t,z) components; cells with u.dsigma < 0 dropped
    void transform_to_local_rest_frame(std::vector<FO_surf> &FOsurf_ptr,
                                       std::vector<FO_surf_LRF> &FOsurf_LRF_ptr);
    void computeFOSurfTmunu(std::vector<FO_surf> &FOsurf_ptr);
    void getParticleQuantumNumbers(long monval, std::array<int, 3> &Qarr);

    // additions of the B200 engine (not in the reference API)
    std::vector<int> read_chosen_particles() const;
    // builds the device sampler (uploads surface, species and tables) without sampling, so that
    // hosts can drive the C ABI (include/iss_cuda.h) on get_sampler()->cuda_handle() themselves
    int prepare_sampler();
    GpuFSSW *get_sampler() { return spectra_sampler_.get(); }
    GpuSpectra *get_spectra() { return efa_.get(); }
    const std::vector<FO_surf> &get_lab_surface() const { return FOsurf_array_; }
    const std::vector<FO_surf_LRF> &get_LRF_surface() const { return FOsurf_LRF_array_; }
    const std::vector<particle_info> &get_particle_table() const { return particle_; }
};

#endif  // ISS_H
