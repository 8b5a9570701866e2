// readindata.h -- freeze-out surface, HRG equation of state and particle-table input.
//
// Host-side ingest for the B200 engine, keeping the file formats and the numerical
// conventions of the reference's read_FOdata (src/readindata.{h,cpp}):
//   * <path>/music_input keys and their precedence over the parameter file
//     (readindata.cpp:25-111),
//   * MUSIC 3+1D and boost-invariant surfaces, text or 34-float binary records
//     (readindata.cpp:395-546, 626-765), cells with T <= 0.01 GeV dropped,
//   * re-derivation of T, mu_B, mu_S, mu_Q, P from the pure-HRG tables, u^0 and q^0
//     renormalisation, transverse-traceless projection of pi^{mu nu}
//     (readindata.cpp:768-842, 1216-1309),
//   * pdg-*.dat particle tables with generated anti-baryons (readindata.cpp:971-1116).
// Surfaces from VISH2+1 / hydro_analysis (hydro_mode 0 and 10) and partial-chemical-
// equilibrium chemical potentials are outside the hot path and are rejected with a message.
#ifndef ISS_B200_READINDATA_H_
#define ISS_B200_READINDATA_H_

#include <cstdint>
#include <string>
#include <vector>

#include "ParameterReader.h"
#include "data_struct.h"

class read_FOdata {
 public:
    read_FOdata(ParameterReader *paraRdr_in, std::string path, std::string table_path,
                std::string particle_table_path);
    ~read_FOdata() { close_surface(); }

    int get_IEOS_music() const { return iEOS_MUSIC_; }
    AfterburnerType get_afterburner_type() const { return afterburner_type_; }
    int get_flag_PCE() const { return flag_PCE_; }
    bool get_surface_in_binary() const { return surface_in_binary_; }

    void read_in_freeze_out_data(std::vector<FO_surf> &surf, std::string surface_filename);
    // Block-wise access to a binary surface (engine addition): the file is loaded once, cells
    // [c0, c0+n) are parsed into `surf` (replacing its content; cells with T <= 0.01 GeV dropped).
    // Lets the facade pipeline parse -> regulate -> LRF transform over cache-sized blocks instead
    // of materialising 160-byte FO_surf records for the whole surface.
    int64_t open_binary_surface(const std::string &surface_filename);   // cells in the file, -1 if text
    void read_binary_block(std::vector<FO_surf> &surf, int64_t c0, int64_t n);
    void close_surface();
    // what the device ingest (iss_cuda_ingest_music_binary) needs from the reader
    const float *binary_records() const { return static_cast<const float *>(surface_map_); }
    bool boost_invariant() const { return mode_ == 1; }
    bool regulates_eos() const {
        return iEOS_MUSIC_ == 9 || iEOS_MUSIC_ == 91 || iEOS_MUSIC_ == 12 || iEOS_MUSIC_ == 14;
    }
    int hrg_nB() const { return (iEOS_MUSIC_ == 12 || iEOS_MUSIC_ == 14) ? 200 : 1; }
    const std::vector<double> &hrg_table() const { return hrg_; }
    void read_in_chemical_potentials(std::vector<FO_surf> &surf,
                                     std::vector<particle_info> &particles);
    int read_resonances_list(std::vector<particle_info> &particles);
    void regulate_surface_cells(std::vector<FO_surf> &surf, bool announce = true);
    void regulate_Wmunu(double u[4], double Wmunu[4][4], double Wmunu_regulated[4][4]);
    int getValuesFromHRGEOS(double ed, double nB, std::vector<double> &eos,
                            std::vector<std::string> *deferred_warnings = nullptr);

 private:
    ParameterReader *paraRdr_;
    const std::string path_, table_path_, particle_table_path_;
    int mode_;
    bool surface_in_binary_;
    bool quantum_statistics_;
    int flag_PCE_;
    int turn_on_bulk_, turn_on_rhob_, turn_on_diff_;
    int iEOS_MUSIC_;
    AfterburnerType afterburner_type_;
    std::vector<double> hrg_;       // rows of 7: ed, nB, P, T, muB, muS, muQ
    void *surface_map_ = nullptr;       // read-only mapping of the binary surface file
    size_t surface_map_bytes_ = 0;
    long hrg_rows_;

    void read_music_input_();
    void read_in_HRG_EOS_();
    void read_binary_surface_(std::vector<FO_surf> &surf, const std::string &file, bool boost_inv);
    void parse_binary_cells_(const float *all, int64_t ncell, bool boost_inv, std::vector<FO_surf> &surf,
                             size_t first);
    void read_text_surface_3d_(std::vector<FO_surf> &surf, const std::string &file);
    void read_text_surface_boost_invariant_(std::vector<FO_surf> &surf, const std::string &file);
};

#endif  // ISS_B200_READINDATA_H_
