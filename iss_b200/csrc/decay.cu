// decay.cu -- resonance feed-down on the device (second kernel of the hot path).
//
// Replaces FSSW::perform_resonance_feed_down (FSSW.cpp:1746-1779) and
// particle_decay::perform_decays / perform_two_body_decay / perform_three_body_decay
// (particle_decay.cpp:265-546).  One thread owns one primary hadron and walks its
// whole decay tree depth-first with a small stack; the walk is run twice with the
// same Philox stream (count pass, then write pass after an integer prefix sum), so
// no intermediate lists are materialised and the output is a pure function of
// (seed, event, index of the primary inside its event).  The count pass walks the tree
// of SPECIES only: which channel is taken depends on the random numbers and on pole
// masses, never on momenta, so it draws the channel, runs the three-body energy
// rejection (the one place where the number of random numbers consumed varies) and
// skips the numbers of the angles and of the life time without generating them.
//
// Reference semantics kept: pole masses (no Breit-Wigner), float mother fields,
// channel pick by cumulative branching ratio, only 2- and 3-body channels emit
// daughters (4-body and N=-2 channels drop the mother, particle_decay.cpp:291-327),
// life time -E/(M Gamma) ln(U) * 0.19733 fm for Gamma > 1e-10.
// Order inside an event: primaries keep their order; each is replaced by its stable
// descendants in depth-first order (the reference appends unstable daughters to the
// end of a work list instead; the multiset of hadrons per event is the same).
#include "iss_internal.cuh"

namespace iss {

constexpr int DECAY_STACK = 24;
constexpr int DECAY_THREADS = 128;
// both passes wait on loads and on their divergent lanes; resident warps hide it better than registers
// do (64 / 48 registers with spills into the L1: 4.96 ms at 4 / 5 CTAs -> 4.62 (5 / 6) -> 4.29 (6 / 8) ->
// 4.13 (8 / 10) -> 4.21 (10 / 12) on the C3 + decays step)
constexpr int DECAY_CTAS_WRITE = 8, DECAY_CTAS_COUNT = 10;

struct DecayArgs {
    const iss_hadron *in;
    int64_t n_in;
    const int64_t *event_off_in;    // [nev+1]
    int64_t nev, ev_begin;
    uint32_t chunk_id;      // surface-chunk mode: first 4096-cell block of the rank (else 0)
    const iss_decay_species *dsp;
    int ndsp;
    const iss_decay_channel *dch;
    const int32_t *sorted_pid;      // [ndsp]
    const int32_t *sorted_idx;      // [ndsp]
    uint64_t seed;
    int64_t *count;                 // [n_in+1] finals per primary (pass 1) / offsets (pass 2)
    iss_hadron *out;
    unsigned long long *errors;     // [0] stack overflow, [1] unknown pid / kinematics
};

struct Part {
    int idx;            // row in dsp
    float mass, E, px, py, pz, t, x, y, z;
};
struct PartSpecies {    // what the count pass keeps of a particle
    int idx;
    float mass;
};
template <bool WRITE> struct PartOf { typedef Part type; };
template <> struct PartOf<false> { typedef PartSpecies type; };

// row of a pid in the particle table: bisection of the CTA's shared-memory copy of the sorted
// (pid, row) pairs (the same search through global memory was a chain of nine dependent loads, the
// first stall of both passes)
__device__ __forceinline__ int find_pid(const int2 *__restrict__ pid_row, int n, int pid) {
    int lo = 0, hi = n - 1;
    while (lo <= hi) {
        const int mid = (lo + hi) >> 1;
        const int2 v = pid_row[mid];
        if (v.x == pid) return v.y;
        if (v.x < pid) lo = mid + 1; else hi = mid - 1;
    }
    return -1;
}

__device__ __forceinline__ void boost_daughter(Part &d, double E_lrf, double px, double py,
                                               double pz, double vx, double vy, double vz,
                                               double v2, double gamma) {
    const double gamma_m_1 = gamma - 1.;
    const double vp = vx*px + vy*py + vz*pz;
    const double f = gamma_m_1*vp/v2 + gamma*E_lrf;
    d.E = static_cast<float>(gamma*(E_lrf + vp));
    d.px = static_cast<float>(px + f*vx);
    d.py = static_cast<float>(py + f*vy);
    d.pz = static_cast<float>(pz + f*vz);
}

template <bool WRITE>
__global__ void __launch_bounds__(DECAY_THREADS, WRITE ? DECAY_CTAS_WRITE : DECAY_CTAS_COUNT)
decay_kernel(const DecayArgs A) {
    extern __shared__ int2 pid_row[];       // [ndsp] sorted by pid
    __shared__ long long ev_first;          // event of the CTA's first primary
    for (int j = threadIdx.x; j < A.ndsp; j += DECAY_THREADS)
        pid_row[j] = make_int2(__ldg(&A.sorted_pid[j]), __ldg(&A.sorted_idx[j]));
    if (threadIdx.x == 0) {
        // one bisection of the event offsets per CTA; its primaries are consecutive, so every
        // thread finds its own event a step or two further on
        const int64_t i0 = static_cast<int64_t>(blockIdx.x)*DECAY_THREADS;
        int64_t lo = 0, hi = A.nev;
        while (hi - lo > 1) {
            const int64_t mid = (lo + hi) >> 1;
            if (__ldg(&A.event_off_in[mid]) <= i0) lo = mid; else hi = mid;
        }
        ev_first = lo;
    }
    __syncthreads();
    const int64_t i = static_cast<int64_t>(blockIdx.x)*DECAY_THREADS + threadIdx.x;
    if (i >= A.n_in) return;
    int64_t ev = ev_first;
    while (ev + 1 < A.nev && __ldg(&A.event_off_in[ev + 1]) <= i) ev++;
    const int64_t k = i - __ldg(&A.event_off_in[ev]);
    Stream rng;
    // (surface-chunk mode: k counts the rank's primaries of the event; the chunk id keeps the
    //  streams of different ranks apart)
    rng.init(A.seed, STREAM_DECAY, A.chunk_id, static_cast<uint32_t>(A.ev_begin + ev),
             static_cast<uint32_t>(k));

    typedef typename PartOf<WRITE>::type P;
    const iss_hadron h0 = A.in[i];
    P stack[DECAY_STACK];
    int sp = 0;
    int64_t nout = 0;
    int64_t wpos = WRITE ? A.count[i] : 0;
    {
        P p;
        p.idx = find_pid(pid_row, A.ndsp, h0.pid);
        p.mass = h0.mass;
        if constexpr (WRITE) {
            p.E = h0.E; p.px = h0.px; p.py = h0.py; p.pz = h0.pz;
            p.t = h0.t; p.x = h0.x; p.y = h0.y; p.z = h0.z;
        }
        if (p.idx < 0) {
            // not in the decay table (cannot happen for species taken from the same pdg file):
            // keep it as it is
            if (WRITE) A.out[wpos] = h0;
            if (!WRITE) A.count[i] = 1;
            return;
        }
        stack[sp++] = p;
    }
    while (sp > 0) {
        const P m = stack[--sp];
        const iss_decay_species ms = A.dsp[m.idx];
        if (ms.stable == 1) {
            if constexpr (WRITE) {
                iss_hadron o;
                o.pid = ms.pid; o.mass = m.mass; o.E = m.E; o.px = m.px; o.py = m.py; o.pz = m.pz;
                o.t = m.t; o.x = m.x; o.y = m.y; o.z = m.z;
                A.out[wpos + nout] = o;
            }
            nout++;
            continue;
        }
        // channel pick (particle_decay.cpp:277-290)
        const double random_local = rng.next();
        double cum = 0.0;
        int pick = ms.first_channel + ms.n_channels - 1;
        for (int c = 0; c < ms.n_channels; c++) {
            cum += A.dch[ms.first_channel + c].branching_ratio;
            if (cum > random_local) {
                pick = ms.first_channel + c;
                break;
            }
        }
        const iss_decay_channel ch = A.dch[pick];
        if (ch.n_part != 2 && ch.n_part != 3) continue;     // mother vanishes (reference quirk)
        bool bad = false;
        for (int d = 0; d < ch.n_part; d++) bad |= (ch.daughter[d] < 0);
        if (bad || sp + ch.n_part > DECAY_STACK) {
            if (WRITE) atomicAdd(&A.errors[bad ? 1 : 0], 1ull);     // (both passes take the same path)
            continue;
        }
        const double M = m.mass;            // float pole mass, as iSS_Hadron stores it
        const double width = ms.width;
        if constexpr (!WRITE) {
            // species and random-number bookkeeping of the same decay (see the header)
            P d[3];
            for (int q = 0; q < ch.n_part; q++) {
                d[q].idx = ch.daughter[q];
                d[q].mass = static_cast<float>(A.dsp[d[q].idx].mass);
            }
            if (ch.n_part == 2) {
                if (M < static_cast<double>(d[0].mass) + static_cast<double>(d[1].mass)) continue;
                rng.skip(2 + (width > 1e-10 ? 1 : 0));      // phi, cos(theta), life time
                stack[sp++] = d[1];
                stack[sp++] = d[0];
            } else {
                const double m1 = d[0].mass, m2 = d[1].mass, m3 = d[2].mass;
                if (M < m1 + m2 + m3) continue;
                const double range = M - m1 - m2 - m3;
                double E1, E2, E3, p1, p2, cos12;
                int guard = 0;
                do {
                    do {
                        E1 = rng.next()*range + m1;
                        E2 = rng.next()*range + m2;
                    } while (E1 + E2 > M);
                    p1 = sqrt(E1*E1 - m1*m1);
                    p2 = sqrt(E2*E2 - m2*m2);
                    E3 = M - E1 - E2;
                    cos12 = (E3*E3 - p1*p1 - p2*p2 - m3*m3)/(2.*p1*p2);
                } while ((cos12 < -1.0 || cos12 > 1.0) && ++guard < 100000);
                rng.skip((width > 1e-10 ? 1 : 0) + 3);      // life time, phi, ksi, cos(theta)
                stack[sp++] = d[2];
                stack[sp++] = d[1];
                stack[sp++] = d[0];
            }
        } else {
        // mother velocity: float divisions (particle_decay.cpp:382-384)
        const double vx = __fdiv_rn(m.px, m.E), vy = __fdiv_rn(m.py, m.E), vz = __fdiv_rn(m.pz, m.E);
        const double v2 = vx*vx + vy*vy + vz*vz;
        const double gamma = 1./sqrt(1. - v2);
        if (ch.n_part == 2) {
            Part d1, d2;
            d1.idx = ch.daughter[0];
            d2.idx = ch.daughter[1];
            d1.mass = static_cast<float>(A.dsp[d1.idx].mass);
            d2.mass = static_cast<float>(A.dsp[d2.idx].mass);
            const double m1 = d1.mass, m2 = d2.mass;
            if (M < m1 + m2) {
                atomicAdd(&A.errors[1], 1ull);
                continue;
            }
            const double temp = M*M - m1*m1 - m2*m2;
            const double p_lrf = sqrt(temp*temp - 4*m1*m1*m2*m2)/(2*M);
            const double phi = rng.next()*2*M_PI;
            const double cos_theta = 2.*(rng.next() - 0.5);
            const double sin_theta = sqrt(1. - cos_theta*cos_theta);
            double sphi, cphi;
            sincos(phi, &sphi, &cphi);
            const double E1 = sqrt(p_lrf*p_lrf + m1*m1);
            const double p1x = p_lrf*sin_theta*cphi, p1y = p_lrf*sin_theta*sphi,
                         p1z = p_lrf*cos_theta;
            const double E2 = sqrt(p_lrf*p_lrf + m2*m2);
            boost_daughter(d1, E1, p1x, p1y, p1z, vx, vy, vz, v2, gamma);
            boost_daughter(d2, E2, -p1x, -p1y, -p1z, vx, vy, vz, v2, gamma);
            double life_time = 1e10;
            if (width > 1e-10) {
                const double tau0 = m.E/(M)*1./(width);
                life_time = -tau0*log(rng.next())*0.19733;
            }
            d1.t = static_cast<float>(m.t + life_time);
            d1.x = static_cast<float>(m.x + __fdiv_rn(m.px, m.E)*life_time);
            d1.y = static_cast<float>(m.y + __fdiv_rn(m.py, m.E)*life_time);
            d1.z = static_cast<float>(m.z + __fdiv_rn(m.pz, m.E)*life_time);
            d2.t = d1.t; d2.x = d1.x; d2.y = d1.y; d2.z = d1.z;
            stack[sp++] = d2;
            stack[sp++] = d1;
        } else {
            Part d1, d2, d3;
            d1.idx = ch.daughter[0];
            d2.idx = ch.daughter[1];
            d3.idx = ch.daughter[2];
            d1.mass = static_cast<float>(A.dsp[d1.idx].mass);
            d2.mass = static_cast<float>(A.dsp[d2.idx].mass);
            d3.mass = static_cast<float>(A.dsp[d3.idx].mass);
            const double m1 = d1.mass, m2 = d2.mass, m3 = d3.mass;
            if (M < m1 + m2 + m3) {
                atomicAdd(&A.errors[1], 1ull);
                continue;
            }
            // (E1, E2, theta12) by accept-reject (particle_decay.cpp:446-458)
            double E1, E2, E3, p1, p2, cos12;
            const double range = M - m1 - m2 - m3;
            int guard = 0;
            do {
                do {
                    E1 = rng.next()*range + m1;
                    E2 = rng.next()*range + m2;
                } while (E1 + E2 > M);
                p1 = sqrt(E1*E1 - m1*m1);
                p2 = sqrt(E2*E2 - m2*m2);
                E3 = M - E1 - E2;
                cos12 = (E3*E3 - p1*p1 - p2*p2 - m3*m3)/(2.*p1*p2);
            } while ((cos12 < -1.0 || cos12 > 1.0) && ++guard < 100000);
            double life_time = 1e10;
            if (width > 1e-10) {
                const double tau = m.E/(M)*1./width;
                life_time = -tau*log(rng.next())*0.19733;
            }
            const float dt = static_cast<float>(m.t + life_time);
            const float dx = static_cast<float>(m.x + __fdiv_rn(m.px, m.E)*life_time);
            const float dy = static_cast<float>(m.y + __fdiv_rn(m.py, m.E)*life_time);
            const float dz = static_cast<float>(m.z + __fdiv_rn(m.pz, m.E)*life_time);
            const double tp2x = p2*sqrt(1. - cos12*cos12);
            const double tp2z = p2*cos12;
            const double tp3x = -tp2x;
            const double tp3z = -(p1 + tp2z);
            const double phi = 2.*M_PI*rng.next();
            const double ksi = 2.*M_PI*rng.next();
            const double cos_theta = 2.*rng.next() - 1.0;
            double sin_phi, cos_phi, sin_ksi, cos_ksi;
            sincos(phi, &sin_phi, &cos_phi);
            sincos(ksi, &sin_ksi, &cos_ksi);
            const double sin_theta = sqrt(1. - cos_theta*cos_theta);
            const double p1x = -p1*sin_theta*cos_ksi;
            const double p1y = p1*sin_theta*sin_ksi;
            const double p1z = p1*cos_theta;
            E1 = sqrt(m1*m1 + p1x*p1x + p1y*p1y + p1z*p1z);
            const double rxx = cos_phi*cos_theta*cos_ksi - sin_phi*sin_ksi;
            const double ryx = -cos_phi*cos_theta*sin_ksi - sin_phi*cos_ksi;
            const double p2x = tp2x*rxx - tp2z*sin_theta*cos_ksi;
            const double p2y = tp2x*ryx + tp2z*sin_theta*sin_ksi;
            const double p2z = tp2x*(cos_phi*sin_theta) + tp2z*cos_theta;
            E2 = sqrt(m2*m2 + p2x*p2x + p2y*p2y + p2z*p2z);
            const double p3x = tp3x*rxx - tp3z*sin_theta*cos_ksi;
            const double p3y = tp3x*ryx + tp3z*(sin_theta*sin_ksi);
            const double p3z = tp3x*cos_phi*sin_theta + tp3z*cos_theta;
            E3 = sqrt(m3*m3 + p3x*p3x + p3y*p3y + p3z*p3z);
            boost_daughter(d1, E1, p1x, p1y, p1z, vx, vy, vz, v2, gamma);
            boost_daughter(d2, E2, p2x, p2y, p2z, vx, vy, vz, v2, gamma);
            boost_daughter(d3, E3, p3x, p3y, p3z, vx, vy, vz, v2, gamma);
            d1.t = dt; d1.x = dx; d1.y = dy; d1.z = dz;
            d2.t = dt; d2.x = dx; d2.y = dy; d2.z = dz;
            d3.t = dt; d3.x = dx; d3.y = dy; d3.z = dz;
            stack[sp++] = d3;
            stack[sp++] = d2;
            stack[sp++] = d1;
        }
        }   // WRITE
    }
    if (!WRITE) A.count[i] = nout;
}

__global__ void gather_event_offsets_kernel(const int64_t *__restrict__ off_primary,
                                            const int64_t *__restrict__ event_off_in, int64_t nev,
                                            int64_t *__restrict__ event_off_out,
                                            int64_t *__restrict__ event_off_host /* mapped */) {
    const int64_t ev = static_cast<int64_t>(blockIdx.x)*blockDim.x + threadIdx.x;
    if (ev <= nev) {
        const int64_t v = off_primary[event_off_in[ev]];
        event_off_out[ev] = v;
        event_off_host[ev] = v;
    }
}

int run_decay(iss_handle *h, uint64_t seed) {
    if (h->decayed) ISS_FAIL(h, ISS_ERR_STATE, "batch already decayed");
    const int64_t n_in = h->n_hadrons;
    const int64_t nev = h->ev_end - h->ev_begin;
    if (n_in == 0) {
        h->decayed = true;
        return ISS_OK;
    }
    int rc = ensure_capacity(h, &h->d_decay_cnt, &h->decay_cnt_cap, n_in + 1 + nev + 1);
    if (rc) return rc;
    if (!h->d_counters) ISS_CUDA_TRY(h, cudaMalloc(&h->d_counters, sizeof(unsigned long long)*N_COUNTERS));
    ISS_CUDA_TRY(h, cudaMemsetAsync(h->d_counters + 4, 0, sizeof(unsigned long long)*2, h->stream));

    DecayArgs A;
    A.in = h->d_hadrons;
    A.n_in = n_in;
    A.event_off_in = h->d_event_off;
    A.nev = nev;
    A.ev_begin = h->ev_begin;
    A.chunk_id = h->chunk ? static_cast<uint32_t>(h->chunk_cell_begin/ISS_CHUNK_ALIGN) : 0u;
    A.dsp = h->d_dsp;
    A.ndsp = h->ndsp;
    A.dch = h->d_dch;
    A.sorted_pid = h->d_sorted_pid;
    A.sorted_idx = h->d_sorted_idx;
    A.seed = seed;
    A.count = h->d_decay_cnt;
    A.out = nullptr;
    A.errors = h->d_counters + 4;

    const unsigned grid = static_cast<unsigned>((n_in + DECAY_THREADS - 1)/DECAY_THREADS);
    const size_t smem = sizeof(int2)*static_cast<size_t>(h->ndsp);
    if (smem > 40*1024) ISS_FAIL(h, ISS_ERR_ARG, "particle table too large for the decay kernel's pid table");
    int64_t total = 0;
    {
        ScopedTimer t(h, ISS_T_DECAY);
        ISS_CUDA_TRY(h, cudaMemsetAsync(h->d_decay_cnt + n_in, 0, sizeof(int64_t), h->stream));
        decay_kernel<false><<<grid, DECAY_THREADS, smem, h->stream>>>(A); ISS_LAUNCHED(h);
        ISS_CUDA_TRY(h, cudaGetLastError());
        rc = device_exclusive_scan_i64(h, h->d_decay_cnt, h->d_decay_cnt, n_in, &total);
        if (rc) return rc;
        if (h->copy_pending2) {
            ISS_CUDA_TRY(h, cudaStreamWaitEvent(h->stream, h->copy_done2, 0));
            h->copy_pending2 = false;
        }
        rc = ensure_capacity(h, &h->d_hadrons2, &h->hadron2_cap, total);
        if (rc) return rc;
        A.out = h->d_hadrons2;
        decay_kernel<true><<<grid, DECAY_THREADS, smem, h->stream>>>(A); ISS_LAUNCHED(h);
        ISS_CUDA_TRY(h, cudaGetLastError());
        // new per-event offsets (into a scratch area behind the counts, then copied over)
        int64_t *tmp = h->d_decay_cnt + n_in + 1;
        gather_event_offsets_kernel<<<static_cast<unsigned>((nev + 1 + 255)/256), 256, 0,
                                      h->stream>>>(h->d_decay_cnt, h->d_event_off, nev, tmp,
                                                    h->d_evoff_mapped); ISS_LAUNCHED(h);
        ISS_CUDA_TRY(h, cudaGetLastError());
        ISS_CUDA_TRY(h, cudaMemcpyAsync(h->d_event_off, tmp, sizeof(int64_t)*(nev + 1),
                                        cudaMemcpyDeviceToDevice, h->stream));
    }
    unsigned long long err[2] = {0, 0};
    rc = mail_post(h, h->d_counters + 4, 2, 16);
    if (rc) return rc;
    ISS_CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    err[0] = *reinterpret_cast<volatile unsigned long long *>(h->h_mail + 16);
    err[1] = *reinterpret_cast<volatile unsigned long long *>(h->h_mail + 17);
    h->n_hadrons = total;
    h->decayed = true;
    if (err[0] || err[1]) {
        char buf[160];
        snprintf(buf, sizeof(buf), "decay: %llu stack overflows, %llu unknown-daughter/kinematic errors",
                 err[0], err[1]);
        ISS_FAIL(h, ISS_ERR_RANGE, buf);
    }
    return ISS_OK;
}

}  // namespace iss
