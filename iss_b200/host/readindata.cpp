// readindata.cpp -- see readindata.h.  Citations are to the reference's src/readindata.cpp.
#include "readindata.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <charconv>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include "logger.h"
#include "parallel.h"

using iSS_data::hbarC;
using iss_host::info;

namespace {

// whole file into memory
bool slurp(const std::string &file, std::vector<char> &buf, bool binary) {
    FILE *f = fopen(file.c_str(), binary ? "rb" : "r");
    if (!f) return false;
    fseek(f, 0, SEEK_END);
    const long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    buf.resize(n + 1);
    const size_t got = fread(buf.data(), 1, n, f);
    fclose(f);
    buf.resize(got + 1);
    buf[got] = '\0';
    return true;
}

// unit conversion and field order of one MUSIC record given 34 numbers (some of which may be
// absent in text files); `direct` holds the values the reference parses straight into float
// members, `viad` the ones that go through a double (`dummy`) first (readindata.cpp:692-749).
struct RawCell {
    float geom[12];     // tau x y eta da0..3 u0..3 (parsed as float)
    double thermo[6];   // e T muB muS muQ (e+P)/T  (parsed as double, fm^-n)
    double pi[10];      // 1/fm^4
    double bulk;        // 1/fm^4
    double rhob;        // 1/fm^3
    float q[4];         // parsed as float
};

inline void convert_cell(const RawCell &r, FO_surf &s) {
    s.tau = r.geom[0]; s.xpt = r.geom[1]; s.ypt = r.geom[2]; s.eta = r.geom[3];
    s.da0 = r.geom[4]; s.da1 = r.geom[5]; s.da2 = r.geom[6]; s.da3 = r.geom[7];
    s.u0 = r.geom[8]; s.u1 = r.geom[9]; s.u2 = r.geom[10]; s.u3 = r.geom[11];
    s.Edec = static_cast<float>(r.thermo[0]*hbarC);
    s.Tdec = static_cast<float>(r.thermo[1]*hbarC);
    s.muB = static_cast<float>(r.thermo[2]*hbarC);
    s.muS = static_cast<float>(r.thermo[3]*hbarC);
    s.muQ = static_cast<float>(r.thermo[4]*hbarC);
    // dummy*Tdec - Edec with the float members promoted (readindata.cpp:718-719)
    s.Pdec = static_cast<float>(r.thermo[5]*s.Tdec - s.Edec);
    float *pi = &s.pi00;
    for (int i = 0; i < 10; i++) pi[i] = static_cast<float>(r.pi[i]*hbarC);
    s.bulkPi = static_cast<float>(r.bulk*hbarC);
    s.Bn = static_cast<float>(r.rhob);
    s.qmu0 = r.q[0]; s.qmu1 = r.q[1]; s.qmu2 = r.q[2]; s.qmu3 = r.q[3];
}

}  // namespace

read_FOdata::read_FOdata(ParameterReader *paraRdr_in, std::string path, std::string table_path,
                         std::string particle_table_path)
    : paraRdr_(paraRdr_in), path_(path), table_path_(table_path),
      particle_table_path_(particle_table_path) {
    mode_ = static_cast<int>(paraRdr_->getVal("hydro_mode"));
    turn_on_bulk_ = static_cast<int>(paraRdr_->getVal("turn_on_bulk"));
    turn_on_rhob_ = static_cast<int>(paraRdr_->getVal("turn_on_rhob"));
    turn_on_diff_ = static_cast<int>(paraRdr_->getVal("turn_on_diff"));
    surface_in_binary_ = true;      // default when music_input has no key (readindata.cpp:35)
    quantum_statistics_ = (paraRdr_->getVal("quantum_statistics") == 1);
    flag_PCE_ = 0;
    iEOS_MUSIC_ = 0;
    hrg_rows_ = 0;
    if (mode_ == 1 || mode_ == 2) {
        info("read in hyper-surface from MUSIC simulations ...");
        read_music_input_();
    } else {
        iss_host::error("hydro_mode 0 (VISH2+1) and 10 (hydro_analysis) surfaces are not supported "
                        "by the B200 engine (MUSIC hydro_mode 1 or 2 only)");
        exit(1);
    }
    const int afterburner_id = static_cast<int>(paraRdr_->getVal("afterburner_type"));
    afterburner_type_ = (afterburner_id == 1)   ? AfterburnerType::UrQMD
                        : (afterburner_id == 2) ? AfterburnerType::SMASH
                                                : AfterburnerType::PDG_Decay;
    // the HRG table decides the particle list (readindata.cpp:104-109)
    if (iEOS_MUSIC_ == 9) afterburner_type_ = AfterburnerType::UrQMD;
    if (iEOS_MUSIC_ == 91) afterburner_type_ = AfterburnerType::SMASH;
    read_in_HRG_EOS_();
}

void read_FOdata::read_music_input_() {
    const std::string file = path_ + "/music_input";
    std::ifstream cfg(file.c_str());
    if (!cfg.is_open()) {
        iss_host::error("read_FOdata::read_FOdata: can not find configuration file " + file);
        exit(1);
    }
    std::string line;
    while (std::getline(cfg, line)) {
        std::stringstream ss(line);
        std::string key;
        ss >> key;
        if (key == "Include_Bulk_Visc_Yes_1_No_0") ss >> turn_on_bulk_;
        else if (key == "Include_Rhob_Yes_1_No_0") ss >> turn_on_rhob_;
        else if (key == "turn_on_baryon_diffusion") ss >> turn_on_diff_;
        else if (key == "EOS_to_use") ss >> iEOS_MUSIC_;
        else if (key == "freeze_surface_in_binary") {
            int flag = 0;
            ss >> flag;
            surface_in_binary_ = (flag == 1);
        }
    }
    if (surface_in_binary_) info("the hyper-surface surface is in the binary format.");
    if (turn_on_bulk_ == 1) info("the hyper-surface includes bulk viscosity.");
    if (turn_on_rhob_ == 1) info("the hyper-surface includes net baryon density.");
    if (turn_on_diff_ == 1) info("the hyper-surface includes baryon diffusion.");
}

// pure-HRG EoS tables (readindata.cpp:913-968).  Only EOS 9/91/12/14 with the UrQMD or SMASH
// list name an existing file; everything else stops here exactly like the reference.
void read_FOdata::read_in_HRG_EOS_() {
    std::cout << " -- Read in pure HRG EoS table...";
    std::string file = table_path_ + "/EOS_tables/";
    if (iEOS_MUSIC_ == 9 || iEOS_MUSIC_ == 91) file += "HRGEOS_PST-";
    else if (iEOS_MUSIC_ == 12) file += "HRGNEOS_B-";
    else if (iEOS_MUSIC_ == 14) file += "HRGNEOS_BQS-";
    if (afterburner_type_ == AfterburnerType::SMASH) file += "SMASH.dat";
    else if (afterburner_type_ == AfterburnerType::UrQMD) file += "urqmd_v3.3+.dat";
    else file = "s95pv1.dat";
    std::vector<char> buf;
    if (!slurp(file, buf, false)) {
        std::cout << "[Error] Can not found EOS file: " << file << std::endl;
        exit(1);
    }
    char *p = buf.data();
    char *eol = strchr(p, '\n');        // header line
    p = eol ? eol + 1 : p + strlen(p);
    hrg_.clear();
    while (*p) {
        eol = strchr(p, '\n');
        if (!eol) break;                // an unterminated last line is not used by the reference
        *eol = '\0';
        double v[7] = {0, 0, 0, 0, 0, 0, 0};
        char *q = p;
        if (iEOS_MUSIC_ == 9 || iEOS_MUSIC_ == 91) {
            // columns: ed P s T  -> {ed, 0, P, T}
            double t[4] = {0, 0, 0, 0};
            for (int i = 0; i < 4; i++) t[i] = strtod(q, &q);
            v[0] = t[0]; v[2] = t[1]; v[3] = t[3];
        } else {
            const int n = (iEOS_MUSIC_ == 12) ? 5 : (iEOS_MUSIC_ == 14 ? 7 : 0);
            for (int i = 0; i < n; i++) v[i] = strtod(q, &q);
        }
        hrg_.insert(hrg_.end(), v, v + 7);
        p = eol + 1;
    }
    hrg_rows_ = static_cast<long>(hrg_.size()/7);
    std::cout << "done." << std::endl;
}

void read_FOdata::read_in_freeze_out_data(std::vector<FO_surf> &surf,
                                          std::string surface_filename) {
    const std::string file = path_ + "/" + surface_filename;
    if (mode_ == 1) {
        std::cout << " -- Read spatial positions of freeze out surface from MUSIC "
                  << "(boost-invariant) ...";
        if (surface_in_binary_) read_binary_surface_(surf, file, true);
        else read_text_surface_boost_invariant_(surf, file);
    } else {
        std::cout << " -- Read spatial positions of freeze out surface from MUSIC...";
        if (surface_in_binary_) read_binary_surface_(surf, file, false);
        else read_text_surface_3d_(surf, file);
    }
    std::cout << "done" << std::endl;
    regulate_surface_cells(surf);
}

// 34 float32 per cell (readindata.cpp:646-689 and 419-461); Pi, rho_B and q^mu are taken from the
// record whatever the turn_on_* flags say.
void read_FOdata::read_binary_surface_(std::vector<FO_surf> &surf, const std::string &file,
                                       bool boost_inv) {
    std::vector<char> buf;
    if (!slurp(file, buf, true)) {
        std::cout << "[Error] Surface file is not found! " << file << std::endl;
        exit(1);
    }
    const size_t ncell = (buf.size() - 1)/(34*sizeof(float));
    parse_binary_cells_(reinterpret_cast<const float *>(buf.data()), static_cast<int64_t>(ncell),
                        boost_inv, surf, surf.size());
}

int64_t read_FOdata::open_binary_surface(const std::string &surface_filename) {
    if (!surface_in_binary_) return -1;
    const std::string file = path_ + "/" + surface_filename;
    close_surface();
    const int fd = open(file.c_str(), O_RDONLY);
    struct stat st;
    if (fd < 0 || fstat(fd, &st) != 0) {
        std::cout << "[Error] Surface file is not found! " << file << std::endl;
        exit(1);
    }
    surface_map_bytes_ = static_cast<size_t>(st.st_size);
    if (surface_map_bytes_ > 0) {
        surface_map_ = mmap(nullptr, surface_map_bytes_, PROT_READ, MAP_PRIVATE, fd, 0);
        if (surface_map_ == MAP_FAILED) {
            surface_map_ = nullptr;
            std::cout << "[Error] can not map surface file " << file << std::endl;
            exit(1);
        }
        madvise(surface_map_, surface_map_bytes_, MADV_SEQUENTIAL);
    }
    close(fd);
    return static_cast<int64_t>(surface_map_bytes_/(34*sizeof(float)));
}

void read_FOdata::read_binary_block(std::vector<FO_surf> &surf, int64_t c0, int64_t n) {
    const float *all = static_cast<const float *>(surface_map_) + 34*c0;
    // the records are overwritten in place: no per-block construction of 160-byte elements
    parse_binary_cells_(all, n, mode_ == 1, surf, 0);
}

void read_FOdata::close_surface() {
    if (surface_map_) munmap(surface_map_, surface_map_bytes_);
    surface_map_ = nullptr;
    surface_map_bytes_ = 0;
}

void read_FOdata::parse_binary_cells_(const float *all, int64_t ncell, bool boost_inv,
                                      std::vector<FO_surf> &surf, size_t first) {
    surf.resize(first + ncell);
    iss_host::parallel_ranges(static_cast<int64_t>(ncell), iss_host::ingest_threads(ncell),
                              [&](int64_t c0, int64_t c1, int) {
        for (int64_t c = c0; c < c1; c++) {
            const float *a = all + 34*c;
            FO_surf &s = surf[first + c];
            s.tau = a[0]; s.xpt = a[1]; s.ypt = a[2];
            s.eta = boost_inv ? 0.0f : a[3];
            s.da0 = a[4]; s.da1 = a[5]; s.da2 = a[6];
            s.da3 = boost_inv ? 0.0f : a[7];
            s.u0 = a[8]; s.u1 = a[9]; s.u2 = a[10]; s.u3 = a[11];
            s.Edec = static_cast<float>(a[12]*hbarC);
            s.Tdec = static_cast<float>(a[13]*hbarC);
            s.muB = static_cast<float>(a[14]*hbarC);
            s.muS = static_cast<float>(a[15]*hbarC);
            s.muQ = static_cast<float>(a[16]*hbarC);
            s.Pdec = a[17]*s.Tdec - s.Edec;         // float arithmetic (readindata.cpp:670)
            float *pi = &s.pi00;
            for (int i = 0; i < 10; i++) pi[i] = static_cast<float>(a[18 + i]*hbarC);
            s.bulkPi = static_cast<float>(a[28]*hbarC);
            s.Bn = a[29];
            s.qmu0 = a[30]; s.qmu1 = a[31]; s.qmu2 = a[32]; s.qmu3 = a[33];
        }
    });
    // cells with T <= 0.01 GeV are discarded, in file order (readindata.cpp:752)
    size_t w = first;
    for (size_t c = first; c < first + static_cast<size_t>(ncell); c++) {
        const FO_surf &s = surf[c];
        if (s.Tdec > 0.01) {
            if (w != c) surf[w] = s;
            w++;
        } else {
            std::cout << "Discard surf elem: T = " << s.Tdec << " GeV, Edec = " << s.Edec
                      << " GeV/fm^3, rhoB = " << s.Bn << " 1/fm^3, muB = " << s.muB << " GeV. "
                      << std::endl;
        }
    }
    surf.resize(w);
}

namespace {

// One number of a text surface.  std::from_chars is correctly rounded like strtof/strtod (same
// bits), needs no locale and is several times faster; it does not skip white space or accept a
// leading '+', and it reports out-of-range values instead of saturating, so those rare tokens go
// through strtof/strtod (the buffer is NUL terminated).
inline void skip_space(const char *&q, const char *end) {
    while (q < end && (*q == ' ' || *q == '\n' || *q == '\t' || *q == '\r' || *q == '\v' || *q == '\f')) q++;
}
inline bool parse_num(const char *&q, const char *end, float &dst) {
    skip_space(q, end);
    if (q >= end) return false;
    const char *t = (*q == '+') ? q + 1 : q;
    const auto r = std::from_chars(t, end, dst);
    if (r.ec == std::errc()) { q = r.ptr; return true; }
    char *e2;
    dst = strtof(q, &e2);
    if (e2 == q || e2 > end) return false;
    q = e2;
    return true;
}
inline bool parse_num(const char *&q, const char *end, double &dst) {
    skip_space(q, end);
    if (q >= end) return false;
    const char *t = (*q == '+') ? q + 1 : q;
    const auto r = std::from_chars(t, end, dst);
    if (r.ec == std::errc()) { q = r.ptr; return true; }
    char *e2;
    dst = strtod(q, &e2);
    if (e2 == q || e2 > end) return false;
    q = e2;
    return true;
}

// [0, len) cut into at most nthread pieces that end just behind a newline
std::vector<size_t> newline_cuts(const char *buf, size_t len, int nthread) {
    std::vector<size_t> cut(1, 0);
    for (int t = 1; t < nthread; t++) {
        size_t pos = len*t/nthread;
        if (pos <= cut.back()) continue;
        const void *nl = memchr(buf + pos, '\n', len - pos);
        if (!nl) break;
        pos = static_cast<const char *>(nl) - buf + 1;
        if (pos > cut.back() && pos < len) cut.push_back(pos);
    }
    cut.push_back(len);
    return cut;
}

struct TextPiece {
    std::vector<FO_surf> cells;     // kept cells of the piece, file order
    std::string messages;           // "Discard surf elem" lines of the piece, file order
    bool partial = false;           // the piece ended inside a cell
};

void discard_message(const FO_surf &s, std::string &out) {
    std::ostringstream os;
    os << "Discard surf elem: T = " << s.Tdec << " GeV, Edec = " << s.Edec
       << " GeV/fm^3, rhoB = " << s.Bn << " 1/fm^3, muB = " << s.muB << " GeV. " << std::endl;
    out += os.str();
}

}  // namespace

// whitespace separated numbers, cells need not be aligned with lines (readindata.cpp:692-749).
// The file is cut at newlines and the pieces are parsed side by side, each assuming that it starts
// at a cell boundary (MUSIC writes one cell per line); if any piece but the last ends inside a
// cell the assumption was wrong and the file is parsed again as one piece.
void read_FOdata::read_text_surface_3d_(std::vector<FO_surf> &surf, const std::string &file) {
    std::vector<char> buf;
    if (!slurp(file, buf, false)) {
        std::cout << "[Error] Surface file is not found! " << file << std::endl;
        exit(1);
    }
    const size_t len = buf.size() - 1;
    const int bulk = turn_on_bulk_, rhob = turn_on_rhob_, diff = turn_on_diff_;
    const char *const file_end = buf.data() + len;
    auto parse_piece = [&](const char *p, const char *end, TextPiece &out) {
        for (;;) {
            RawCell r;
            memset(&r, 0, sizeof(r));
            const char *q = p;
            bool ok = true;
            int got = 0;
            for (int i = 0; i < 12 && ok; i++) { ok = parse_num(q, end, r.geom[i]); got += ok; }
            for (int i = 0; i < 6 && ok; i++) { ok = parse_num(q, end, r.thermo[i]); got += ok; }
            for (int i = 0; i < 10 && ok; i++) { ok = parse_num(q, end, r.pi[i]); got += ok; }
            if (bulk == 1 && ok) { ok = parse_num(q, end, r.bulk); got += ok; }
            if (rhob == 1 && ok) { ok = parse_num(q, end, r.rhob); got += ok; }
            if (diff == 1)
                for (int i = 0; i < 4 && ok; i++) { ok = parse_num(q, end, r.q[i]); got += ok; }
            if (!ok) {          // end of data (the reference stops at stream eof)
                out.partial = got > 0;
                break;
            }
            if (q >= file_end) {
                // the cell's last number ends exactly at the end of the file: the reference's
                // stream extraction has set eofbit by then and the cell is NOT kept
                // (`if (!surfdat.eof())`, readindata.cpp:752)
                out.partial = true;
                break;
            }
            p = q;
            FO_surf s;
            convert_cell(r, s);
            if (s.Tdec > 0.01) out.cells.push_back(s);
            else discard_message(s, out.messages);
        }
    };
    int nthread = iss_host::ingest_threads(static_cast<int64_t>(len), 1 << 20);
    std::vector<TextPiece> pieces;
    for (int attempt = 0; attempt < 2; attempt++) {
        const std::vector<size_t> cut = newline_cuts(buf.data(), len, nthread);
        const int np = static_cast<int>(cut.size()) - 1;
        pieces.assign(np, TextPiece());
        iss_host::parallel_ranges(np, np, [&](int64_t b, int64_t e, int) {
            for (int64_t k = b; k < e; k++) {
                pieces[k].cells.reserve((cut[k + 1] - cut[k])/400 + 16);
                parse_piece(buf.data() + cut[k], buf.data() + cut[k + 1], pieces[k]);
            }
        });
        bool aligned = true;
        for (int k = 0; k + 1 < np; k++) aligned = aligned && !pieces[k].partial;
        if (aligned) break;
        nthread = 1;            // cells straddle lines: one piece, the reference's stream order
    }
    size_t total = surf.size();
    for (const TextPiece &pc : pieces) total += pc.cells.size();
    surf.reserve(total);
    for (const TextPiece &pc : pieces) {
        if (!pc.messages.empty()) std::cout << pc.messages;
        surf.insert(surf.end(), pc.cells.begin(), pc.cells.end());
    }
}

// one cell per line, eta and da3 forced to zero (readindata.cpp:464-529); lines are independent,
// so the pieces are parsed side by side
void read_FOdata::read_text_surface_boost_invariant_(std::vector<FO_surf> &surf,
                                                     const std::string &file) {
    std::vector<char> buf;
    if (!slurp(file, buf, false)) {
        std::cout << "[Error] Surface file is not found! " << file << std::endl;
        exit(1);
    }
    // the reference only keeps a line if the stream is not at eof after reading it, i.e. the
    // line is terminated by a newline (readindata.cpp:531-541)
    size_t len = buf.size() - 1;
    while (len > 0 && buf[len - 1] != '\n') len--;
    const int bulk = turn_on_bulk_, rhob = turn_on_rhob_, diff = turn_on_diff_;
    auto parse_piece = [&](const char *p, const char *end, TextPiece &out) {
        while (p < end) {
            const char *eol = static_cast<const char *>(memchr(p, '\n', end - p));
            if (!eol) break;
            RawCell r;
            memset(&r, 0, sizeof(r));
            const char *q = p;
            // a short line leaves the remaining fields at zero, like strtod on an exhausted string
            double d4[4] = {0., 0., 0., 0.};
            for (int i = 0; i < 4; i++) parse_num(q, eol, d4[i]);       // tau x y eta via doubles
            for (int i = 0; i < 4; i++) r.geom[i] = static_cast<float>(d4[i]);
            for (int i = 4; i < 12; i++) parse_num(q, eol, r.geom[i]);
            for (int i = 0; i < 6; i++) parse_num(q, eol, r.thermo[i]);
            for (int i = 0; i < 10; i++) parse_num(q, eol, r.pi[i]);
            if (bulk == 1) parse_num(q, eol, r.bulk);
            if (rhob == 1) parse_num(q, eol, r.rhob);
            if (diff == 1)
                for (int i = 0; i < 4; i++) parse_num(q, eol, r.q[i]);
            r.geom[3] = 0.0f;       // eta
            r.geom[7] = 0.0f;       // da3
            FO_surf s;
            convert_cell(r, s);
            if (s.Tdec > 0.01) out.cells.push_back(s);
            else discard_message(s, out.messages);
            p = eol + 1;
        }
    };
    const int nthread = iss_host::ingest_threads(static_cast<int64_t>(len), 1 << 20);
    const std::vector<size_t> cut = newline_cuts(buf.data(), len, nthread);
    const int np = static_cast<int>(cut.size()) - 1;
    std::vector<TextPiece> pieces(np);
    iss_host::parallel_ranges(np, np, [&](int64_t b, int64_t e, int) {
        for (int64_t k = b; k < e; k++) {
            pieces[k].cells.reserve((cut[k + 1] - cut[k])/400 + 16);
            parse_piece(buf.data() + cut[k], buf.data() + cut[k + 1], pieces[k]);
        }
    });
    size_t total = surf.size();
    for (const TextPiece &pc : pieces) total += pc.cells.size();
    surf.reserve(total);
    for (const TextPiece &pc : pieces) {
        if (!pc.messages.empty()) std::cout << pc.messages;
        surf.insert(surf.end(), pc.cells.begin(), pc.cells.end());
    }
}

// readindata.cpp:768-842
void read_FOdata::regulate_surface_cells(std::vector<FO_surf> &surf, bool announce) {
    const bool regulateTemperature =
        (iEOS_MUSIC_ == 9 || iEOS_MUSIC_ == 91 || iEOS_MUSIC_ == 12 || iEOS_MUSIC_ == 14);
    if (regulateTemperature && announce)
        std::cout << "Regulate local temperature with pure HRG EoS." << std::endl;
    const int64_t n = static_cast<int64_t>(surf.size());
    // "ed is out of range" warnings are collected per thread and printed in cell order afterwards
    // (the reference is serial: its warnings come in file order)
    const int nthread = iss_host::ingest_threads(n);
    std::vector<std::vector<std::string>> warnings(std::max(1, nthread));
    iss_host::parallel_ranges(n, nthread, [&](int64_t c0, int64_t c1, int tix) {
        std::vector<double> eos;
        for (int64_t c = c0; c < c1; c++) {
            FO_surf &s = surf[c];
            if (regulateTemperature) {
                if (getValuesFromHRGEOS(s.Edec, s.Bn, eos, &warnings[tix]) == 0) {
                    s.Tdec = static_cast<float>(eos[1]);
                    s.muB = static_cast<float>(eos[2]);
                    s.muS = static_cast<float>(eos[3]);
                    s.muQ = static_cast<float>(eos[4]);
                    s.Pdec = static_cast<float>(eos[0]);
                }
            }
            // 1. is a double literal, the products are float (readindata.cpp:796-798)
            s.u0 = static_cast<float>(std::sqrt(1. + s.u1*s.u1 + s.u2*s.u2 + s.u3*s.u3));
            s.qmu0 = (s.u1*s.qmu1 + s.u2*s.qmu2 + s.u3*s.qmu3)/s.u0;
            double u[4] = {s.u0, s.u1, s.u2, s.u3};
            double W[4][4] = {{s.pi00, s.pi01, s.pi02, s.pi03},
                              {s.pi01, s.pi11, s.pi12, s.pi13},
                              {s.pi02, s.pi12, s.pi22, s.pi23},
                              {s.pi03, s.pi13, s.pi23, s.pi33}};
            double R[4][4];
            regulate_Wmunu(u, W, R);
            s.pi00 = static_cast<float>(R[0][0]); s.pi01 = static_cast<float>(R[0][1]);
            s.pi02 = static_cast<float>(R[0][2]); s.pi03 = static_cast<float>(R[0][3]);
            s.pi11 = static_cast<float>(R[1][1]); s.pi12 = static_cast<float>(R[1][2]);
            s.pi13 = static_cast<float>(R[1][3]); s.pi22 = static_cast<float>(R[2][2]);
            s.pi23 = static_cast<float>(R[2][3]); s.pi33 = static_cast<float>(R[3][3]);
        }
    });
    for (const auto &per_thread : warnings)
        for (const std::string &w : per_thread) iss_host::warning(w);
}

// transverse, traceless projection (readindata.cpp:1216-1246)
void read_FOdata::regulate_Wmunu(double u[4], double W[4][4], double R[4][4]) {
    const double g[4] = {-1., 1., 1., 1.};
    double u_dot_pi[4], u_mu[4];
    for (int i = 0; i < 4; i++) {
        u_dot_pi[i] = -u[0]*W[0][i] + u[1]*W[1][i] + u[2]*W[2][i] + u[3]*W[3][i];
        u_mu[i] = g[i]*u[i];
    }
    const double tr_pi = -W[0][0] + W[1][1] + W[2][2] + W[3][3];
    double upu = 0.0;
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) upu += u_mu[i]*W[i][j]*u_mu[j];
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) {
            const double gij = (i == j) ? g[i] : 0.;
            R[i][j] = (W[i][j] + u[i]*u_dot_pi[j] + u[j]*u_dot_pi[i] + u[i]*u[j]*upu
                       - 1./3.*(gij + u[i]*u[j])*(tr_pi + upu));
        }
}

// bilinear (e, n_B) interpolation of the HRG table (readindata.cpp:1249-1309)
int read_FOdata::getValuesFromHRGEOS(double ed, double nB, std::vector<double> &eosVar,
                                     std::vector<std::string> *deferred_warnings) {
    eosVar.assign(5, 0.);       // {P, T, muB, muS, muQ}
    const int nBlen = (iEOS_MUSIC_ == 12 || iEOS_MUSIC_ == 14) ? 200 : 1;
    auto H = [&](long row, int col) { return hrg_[row*7 + col]; };
    const double de = H(nBlen, 0) - H(0, 0);
    const double e0 = H(0, 0);
    const int e_idx = static_cast<int>((ed - e0)/de);
    if (e_idx < 0 || e_idx >= static_cast<int>(hrg_rows_/nBlen) - 2) {
        std::ostringstream os;
        os << "ed is out of range: ed = " << ed << " GeV/fm^3. Can not regulate this fluid cell!";
        if (deferred_warnings) deferred_warnings->push_back(os.str());
        else iss_host::warning(os.str());
        return -1;
    }
    const long r1 = static_cast<long>(e_idx)*nBlen;
    const long r2 = static_cast<long>(e_idx + 1)*nBlen;
    const double e_frac = (ed - H(r1, 0))/de;
    double f1 = 0, f2 = 0;
    int i1 = 0, i2 = 0;
    if (nBlen > 1) {
        const double dnB1 = H(r1 + 1, 1), dnB2 = H(r2 + 1, 1);
        i1 = std::min(nBlen - 2, static_cast<int>(nB/dnB1));
        i2 = std::min(nBlen - 2, static_cast<int>(nB/dnB2));
        f1 = std::min(1., (nB - H(r1 + i1, 1))/dnB1);
        f2 = std::min(1., (nB - H(r2 + i2, 1))/dnB2);
    }
    auto interp = [&](int col) {
        const double a = H(r1 + i1, col)*(1. - f1) + H(r1 + i1 + 1, col)*f1;
        const double b = H(r2 + i2, col)*(1. - f2) + H(r2 + i2 + 1, col)*f2;
        return a*(1 - e_frac) + b*e_frac;
    };
    eosVar[0] = interp(2);
    eosVar[1] = interp(3);
    if (nBlen > 1)
        for (int c = 4; c < 7; c++) eosVar[c - 2] = interp(c);
    return 0;
}

void read_FOdata::read_in_chemical_potentials(std::vector<FO_surf> &surf,
                                              std::vector<particle_info> &particles) {
    (void)surf;
    // readindata.cpp:195-300: every EOS for which the HRG table exists (9, 91, 12, 14) is in
    // chemical equilibrium, N_stableparticle = 0.
    const int Nparticle = read_resonances_list(particles);
    std::ostringstream os;
    os << "total number of particle species: " << Nparticle;
    info(os.str());
    info(" -- EOS is chemical equilibrium. ");
    flag_PCE_ = 0;
}

// pdg-*.dat: "monval name mass width gspin baryon strange charm bottom gisospin charge decays"
// followed by `decays` lines "monval Npart BR d1 d2 d3 d4 d5"; anti-baryons are generated with
// conjugated daughters (readindata.cpp:971-1116).
int read_FOdata::read_resonances_list(std::vector<particle_info> &particle) {
    const double eps = 1e-15;
    std::cout << " -- Read in particle resonance decay table...";
    std::string file;
    if (afterburner_type_ == AfterburnerType::SMASH) file = particle_table_path_ + "/pdg-SMASH.dat";
    else if (afterburner_type_ == AfterburnerType::UrQMD)
        file = particle_table_path_ + "/pdg-urqmd_v3.3+.dat";
    else file = particle_table_path_ + "/pdg-s95pv1.dat";
    std::ifstream in(file.c_str());
    if (!in.good()) {
        std::cout << "[Error] Can not found pdg file: " << file << std::endl;
        exit(1);
    }
    for (;;) {
        particle_info p;
        if (!(in >> p.monval)) break;
        in >> p.name >> p.mass >> p.width >> p.gspin >> p.baryon >> p.strange >> p.charm
           >> p.bottom >> p.gisospin >> p.charge >> p.decays;
        for (int j = 0; j < p.decays; j++) {
            auto *ch = new decay_channel_info;
            int own;
            in >> own >> ch->decay_Npart >> ch->branching_ratio;
            for (int k = 0; k < 5; k++) in >> ch->decay_part[k];
            std::string rest;
            std::getline(in, rest);
            p.decay_channels.push_back(ch);
        }
        p.stable = (!p.decay_channels.empty() && p.decay_channels[0]->decay_Npart == 1) ? 1 : 0;
        p.sign = 0;
        particle.push_back(p);
        if (p.baryon > 0) {
            particle_info a = p;
            a.monval = -p.monval;
            a.name = "Anti-" + p.name;
            a.baryon = -p.baryon;
            a.strange = -p.strange;
            a.charm = -p.charm;
            a.bottom = -p.bottom;
            a.charge = -p.charge;
            a.decay_channels.clear();
            for (int j = 0; j < p.decays; j++) {
                auto *ch = new decay_channel_info(*p.decay_channels[j]);
                for (int k = 0; k < 5; k++) {
                    const int d = p.decay_channels[j]->decay_part[k];
                    if (d == 0) continue;
                    size_t idx = 0;
                    // search among the entries read so far (baryon included, its own
                    // anti-particle not yet)
                    const size_t nsearch = particle.size();
                    for (; idx < nsearch; idx++)
                        if (particle[idx].monval == d) break;
                    if (idx == nsearch) {
                        if (p.stable == 0 && ch->branching_ratio > eps) {
                            iss_host::error("Can not find decay particle index for anti-baryon!");
                            std::ostringstream os;
                            os << "particle monval : " << d;
                            iss_host::error(os.str());
                            exit(1);
                        }
                        ch->decay_part[k] = -d;
                        continue;
                    }
                    const particle_info &dp = particle[idx];
                    ch->decay_part[k] = (dp.baryon == 0 && dp.charge == 0 && dp.strange == 0) ? d : -d;
                }
                a.decay_channels.push_back(ch);
            }
            particle.push_back(a);
        }
    }
    for (auto &p : particle)
        p.sign = quantum_statistics_ ? (p.baryon == 0 ? -1 : 1) : 0;
    std::cout << "done." << std::endl;
    return static_cast<int>(particle.size());
}
