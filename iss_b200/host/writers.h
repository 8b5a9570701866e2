// writers.h -- the reference's three sample-file formats (FSSW.cpp:365-561), as free functions
// over a contiguous hadron buffer + event offsets.  Files are byte-compatible with the
// reference's (tests/test_writers_cpu.py compares against files written by the reference).
#ifndef ISS_B200_WRITERS_H_
#define ISS_B200_WRITERS_H_

#include <cstdint>
#include <string>

#include "data_struct.h"

namespace iss_writers {

// OSCAR1997A text: header file verbatim, then per NON-EMPTY event "iev(0-based) N 0 0" and per
// hadron "index pid" + px py pz E m x y z t in %24.16e (FSSW.cpp:365-494)
void write_oscar(const std::string &filename, const std::string &header_file,
                 const iSS_Hadron *hadrons, const int64_t *event_off, int64_t nev);
// "N \n" per event, then "pid " + mass t x y z E px py pz in %.7e, gzip-compressed
// (FSSW.cpp:497-526)
void write_gzip(const std::string &filename, const iSS_Hadron *hadrons, const int64_t *event_off,
                int64_t nev);
// int N per event, then per hadron int pid + 9 float {mass,t,x,y,z,E,px,py,pz} (FSSW.cpp:529-561)
void write_binary(const std::string &filename, const iSS_Hadron *hadrons, const int64_t *event_off,
                  int64_t nev);

}  // namespace iss_writers
#endif  // ISS_B200_WRITERS_H_
