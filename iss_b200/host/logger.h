// logger.h -- "[Info] <max RSS> MB message" console lines in the style of the reference's
// pretty_ostream (src/pretty_ostream.cpp:35-76).  Cosmetic; kept so logs stay greppable.
#ifndef ISS_B200_LOGGER_H_
#define ISS_B200_LOGGER_H_

#include <sys/resource.h>

#include <iomanip>
#include <iostream>
#include <sstream>
#include <string>

namespace iss_host {

inline std::string rss_mb() {
    struct rusage u;
    std::ostringstream os;
    if (getrusage(RUSAGE_SELF, &u) == 0) os << std::setprecision(4) << u.ru_maxrss/1024. << " MB";
    return os.str();
}
inline void info(const std::string &m) { std::cout << "[Info] " << rss_mb() << " " << m << std::endl; }
inline void warning(const std::string &m) { std::cout << "\033[1m\033[33m[Warning] " << m << "\033[0m" << std::endl; }
inline void error(const std::string &m) { std::cout << "\033[1m\033[31m[Error] " << m << "\033[0m" << std::endl; }

}  // namespace iss_host
#endif  // ISS_B200_LOGGER_H_
