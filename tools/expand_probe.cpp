// expand_probe.cpp -- host-side cost of expanding 20-byte wire records (cell, px, py, pz, E) into 40-byte
// iSS_Hadron records (species runs of ~170, cell look-up in a 16 MB table): decides whether halving
// the PCIe bytes pays against the device->host DMA (~53 GB/s).  g++ -O3 -march=native -lpthread.
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
#include <immintrin.h>
struct Wire { uint32_t cell; float px, py, pz, E; };
struct Had { int32_t pid; float mass, E, px, py, pz, t, x, y, z; };
static void expand(const Wire *w, Had *o, int64_t n, const float *txyz, int32_t pid, float mass) {
    for (int64_t i = 0; i < n; i++) {
        const Wire r = w[i];
        const float *c = txyz + 4ull*r.cell;
        Had h;
        h.pid = pid; h.mass = mass; h.E = r.E; h.px = r.px; h.py = r.py; h.pz = r.pz;
        h.t = c[0]; h.x = c[1]; h.y = c[2]; h.z = c[3];
        o[i] = h;
    }
}
int main(int argc, char **argv) {
    const int nt = argc > 1 ? atoi(argv[1]) : 8;
    const int64_t n = 54800000, ncell = 1000000;
    std::vector<Wire> w(n);
    std::vector<Had> o(n);
    std::vector<float> txyz(4*ncell, 1.f);
    uint64_t s = 88172645463325252ull;
    for (int64_t i = 0; i < n; i++) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; w[i].cell = s % ncell; w[i].px = 1; w[i].py = 2; w[i].pz = 3; w[i].E = 4; }
    memset(o.data(), 0, n*sizeof(Had));
    for (int rep = 0; rep < 3; rep++) {
        auto t0 = std::chrono::steady_clock::now();
        std::vector<std::thread> th;
        for (int t = 0; t < nt; t++) th.emplace_back([&, t] {
            const int64_t b = n*t/nt, e = n*(t + 1)/nt;
            // runs of ~170 hadrons per (event, species)
            for (int64_t i = b; i < e; i += 170) expand(&w[i], &o[i], std::min<int64_t>(170, e - i), txyz.data(), 211 + (int)(i & 7), 0.139f);
        });
        for (auto &x : th) x.join();
        double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        printf("threads %d: %.1f ms, %.2f G rec/s, %.1f GB/s (60 B/rec)\n", nt, dt*1e3, n/dt/1e9, n*60.0/dt/1e9);
    }
}
