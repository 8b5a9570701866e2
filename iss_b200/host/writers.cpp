// writers.cpp -- see writers.h.  Citations are to the reference's src/FSSW.cpp.
#include "writers.h"

#include <zlib.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <string>
#include <thread>
#include <vector>

#include "logger.h"

namespace {

double seconds_since(const std::chrono::steady_clock::time_point &t0) {
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

// "%10d  %10d  " + nine %24.16e separated by two blanks + newline: what `oscar << setw(10) << i
// << "  " << setw(10) << pid << "  " << line_buffer << endl` produces (FSSW.cpp:468-483)
inline int format_oscar_line(char *dst, size_t cap, long index, const iSS_Hadron &hd) {
    return snprintf(dst, cap,
                    "%10ld  %10d  %24.16e  %24.16e  %24.16e  %24.16e  %24.16e  %24.16e  %24.16e  "
                    "%24.16e  %24.16e\n",
                    index, hd.pid, hd.px, hd.py, hd.pz, hd.E, hd.mass, hd.x, hd.y, hd.z, hd.t);
}

}  // namespace

namespace iss_writers {

void write_oscar(const std::string &filename, const std::string &header_file,
                 const iSS_Hadron *hadrons, const int64_t *event_off, int64_t nev) {
    const auto t0 = std::chrono::steady_clock::now();
    iss_host::info(" -- Now combine sample files to OSCAR file...");
    remove(filename.c_str());
    std::ifstream header(header_file.c_str());
    if (!header.is_open()) {
        std::cout << std::endl
                  << "combine_samples_to_OSCAR error: OSCAR header file " << header_file
                  << " not found." << std::endl;
        exit(-1);
    }
    FILE *out = fopen(filename.c_str(), "w");
    if (!out) {
        iss_host::error("can not open " + filename);
        exit(-1);
    }
    std::string line;
    while (std::getline(header, line)) {
        if (header.eof()) break;    // the reference drops an unterminated last line (:383-390)
        fputs(line.c_str(), out);
        fputc('\n', out);
    }
    // The text of one event is formatted by several threads into per-thread buffers (the
    // %24.16e conversions dominate: ~260 bytes per hadron) and written in order.
    constexpr int LINE = 280;
    const int nthread = static_cast<int>(std::max(1u, std::min(16u, std::thread::hardware_concurrency())));
    std::vector<std::vector<char>> bufs(nthread);
    for (int64_t ev = 0; ev < nev; ev++) {
        const int64_t n = event_off[ev + 1] - event_off[ev];
        if (n <= 0) continue;       // empty events are skipped (:461-462)
        // setw(8) << 0.0 prints "0" in a field of 8
        fprintf(out, "%10ld  %10ld  %8s  %8s\n", static_cast<long>(ev), static_cast<long>(n), "0", "0");
        const iSS_Hadron *h = hadrons + event_off[ev];
        const int use = static_cast<int>(std::min<int64_t>(nthread, (n + 4095)/4096));
        auto work = [&](int t) {
            const int64_t a = n*t/use, b = n*(t + 1)/use;
            std::vector<char> &buf = bufs[t];
            buf.resize(static_cast<size_t>(b - a)*LINE + 1);
            size_t pos = 0;
            for (int64_t i = a; i < b; i++)
                pos += format_oscar_line(buf.data() + pos, LINE + 1, static_cast<long>(i + 1), h[i]);
            buf.resize(pos);
        };
        if (use <= 1) {
            work(0);
        } else {
            std::vector<std::thread> pool;
            for (int t = 0; t < use; t++) pool.emplace_back(work, t);
            for (auto &t : pool) t.join();
        }
        for (int t = 0; t < use; t++) fwrite(bufs[t].data(), 1, bufs[t].size(), out);
    }
    fclose(out);
    std::cout << std::endl
              << " -- combine_samples_to_OSCAR samples finishes " << seconds_since(t0)
              << " seconds." << std::endl;
}

void write_gzip(const std::string &filename, const iSS_Hadron *hadrons, const int64_t *event_off,
                int64_t nev) {
    const auto t0 = std::chrono::steady_clock::now();
    iss_host::info(" -- Now combine sample files to a gzip file...");
    remove(filename.c_str());
    gzFile fp = gzopen(filename.c_str(), "wb");
    for (int64_t ev = 0; ev < nev; ev++) {
        const int n = static_cast<int>(event_off[ev + 1] - event_off[ev]);
        gzprintf(fp, "%d \n", n);
        for (int64_t i = event_off[ev]; i < event_off[ev + 1]; i++) {
            const iSS_Hadron &hd = hadrons[i];
            gzprintf(fp, "%d ", hd.pid);
            gzprintf(fp, "%.7e %.7e %.7e %.7e %.7e %.7e %.7e %.7e %.7e\n", hd.mass, hd.t, hd.x,
                     hd.y, hd.z, hd.E, hd.px, hd.py, hd.pz);
        }
    }
    gzclose(fp);
    std::cout << std::endl
              << " -- combine_samples_to_gzip_file finishes " << seconds_since(t0) << " seconds."
              << std::endl;
}

void write_binary(const std::string &filename, const iSS_Hadron *hadrons, const int64_t *event_off,
                  int64_t nev) {
    const auto t0 = std::chrono::steady_clock::now();
    iss_host::info(" -- Now combine sample files to a binary file...");
    remove(filename.c_str());
    FILE *out = fopen(filename.c_str(), "wb");
    if (!out) {
        iss_host::error("can not open " + filename);
        exit(-1);
    }
    std::vector<char> rec;
    for (int64_t ev = 0; ev < nev; ev++) {
        const int n = static_cast<int>(event_off[ev + 1] - event_off[ev]);
        rec.resize(sizeof(int) + static_cast<size_t>(n)*40);
        char *p = rec.data();
        memcpy(p, &n, sizeof(int));
        p += sizeof(int);
        for (int64_t i = event_off[ev]; i < event_off[ev + 1]; i++) {
            const iSS_Hadron &hd = hadrons[i];
            const float a[9] = {hd.mass, hd.t, hd.x, hd.y, hd.z, hd.E, hd.px, hd.py, hd.pz};
            memcpy(p, &hd.pid, sizeof(int));
            memcpy(p + sizeof(int), a, sizeof(a));
            p += 40;
        }
        fwrite(rec.data(), 1, rec.size(), out);
    }
    fclose(out);
    std::cout << std::endl
              << " -- combine_samples_to_binary_file finishes " << seconds_since(t0) << " seconds."
              << std::endl;
}

}  // namespace iss_writers
