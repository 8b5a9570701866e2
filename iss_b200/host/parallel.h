// parallel.h -- minimal fork-join helper for the host-side ingest loops (one std::thread per
// contiguous range; the per-cell arithmetic itself is unchanged, so results do not depend on the
// number of threads).
#ifndef ISS_B200_PARALLEL_H_
#define ISS_B200_PARALLEL_H_

#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <thread>
#include <vector>

namespace iss_host {

inline int ingest_threads(int64_t n, int64_t min_per_thread = 16384) {
    int hw = static_cast<int>(std::thread::hardware_concurrency());
    if (const char *e = getenv("ISS_HOST_THREADS")) hw = atoi(e);
    hw = std::max(1, std::min(hw, 32));
    const int64_t by_size = std::max<int64_t>(1, n/min_per_thread);
    return static_cast<int>(std::min<int64_t>(hw, by_size));
}

// fn(begin, end, thread_index) over [0, n) split into nthread contiguous ranges
template <typename F>
void parallel_ranges(int64_t n, int nthread, F fn) {
    if (nthread <= 1 || n <= 0) {
        fn(static_cast<int64_t>(0), n, 0);
        return;
    }
    std::vector<std::thread> pool;
    pool.reserve(nthread);
    for (int t = 0; t < nthread; t++)
        pool.emplace_back(fn, n*t/nthread, n*(t + 1)/nthread, t);
    for (auto &th : pool) th.join();
}

}  // namespace iss_host
#endif  // ISS_B200_PARALLEL_H_
