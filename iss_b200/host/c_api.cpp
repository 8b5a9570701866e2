// c_api.cpp -- extern "C" forwarding layer over class iSS; see include/iss_host.h.
#include <cstring>
#include <string>

#include "../../include/iss_host.h"
#include "gpu_fssw.h"
#include "gpu_spectra.h"
#include "iSS.h"
#include "writers.h"

struct iss_host {
    iSS *obj;
};

static_assert(sizeof(iss_hadron) == sizeof(iSS_Hadron), "record layouts must agree");

extern "C" {

iss_host *iss_host_create(const char *path, const char *table_path,
                          const char *particle_table_path, const char *inputfile,
                          const char *surface_filename) {
    iss_host *s = new iss_host;
    s->obj = new iSS(path, table_path ? table_path : "iSS_tables",
                     particle_table_path ? particle_table_path : "iSS_tables",
                     inputfile ? inputfile : "iSS_parameters.dat",
                     surface_filename ? surface_filename : "surface.dat");
    return s;
}

void iss_host_destroy(iss_host *s) {
    if (!s) return;
    delete s->obj;
    delete s;
}

void iss_host_set_param(iss_host *s, const char *name, double value) {
    s->obj->paraRdr_ptr->setVal(name, value);
}

double iss_host_get_param(iss_host *s, const char *name, double default_value) {
    if (!s->obj->paraRdr_ptr->exist(name)) return default_value;
    return s->obj->paraRdr_ptr->getVal(name);
}

void iss_host_parse_param(iss_host *s, const char *kv) { s->obj->paraRdr_ptr->phraseOneLine(kv); }
void iss_host_set_random_seed(iss_host *s, int seed) { s->obj->set_random_seed(seed); }
int iss_host_read_in_FO_surface(iss_host *s) { return s->obj->read_in_FO_surface(); }
int iss_host_generate_samples(iss_host *s) { return s->obj->generate_samples(); }
int iss_host_shell(iss_host *s) { return s->obj->shell(); }
void iss_host_perform_checks(iss_host *s) { s->obj->perform_checks(); }
int iss_host_get_number_of_sampled_events(iss_host *s) { return s->obj->get_number_of_sampled_events(); }
int iss_host_get_number_of_particles(iss_host *s, int iev) { return s->obj->get_number_of_particles(iev); }

const iss_hadron *iss_host_get_hadron_list_iev(iss_host *s, int iev, int64_t *n) {
    std::vector<iSS_Hadron> *v = s->obj->get_hadron_list_iev(iev);
    if (n) *n = static_cast<int64_t>(v->size());
    return reinterpret_cast<const iss_hadron *>(v->data());
}

void iss_host_clear(iss_host *s) { s->obj->clear(); }
int iss_host_prepare_sampler(iss_host *s) { return s->obj->prepare_sampler(); }

iss_handle *iss_host_cuda_handle(iss_host *s) {
    GpuFSSW *g = s->obj->get_sampler();
    return g ? g->cuda_handle() : nullptr;
}

int64_t iss_host_lrf_surface(iss_host *s, float *dst) {
    const auto &surf = s->obj->get_LRF_surface();
    if (dst) {
        for (const auto &c : surf) {
            const float rec[ISS_NFIELD] = {
                c.tau, c.xpt, c.ypt, c.eta,
                c.da_mu_LRF[0], c.da_mu_LRF[1], c.da_mu_LRF[2], c.da_mu_LRF[3],
                c.u_tz[0], c.u_tz[1], c.u_tz[2], c.u_tz[3],
                c.Edec, c.Tdec, c.Pdec, c.Bn, c.muB, c.muS, c.muQ, c.bulkPi,
                c.piLRF_xx, c.piLRF_xy, c.piLRF_xz, c.piLRF_yy, c.piLRF_yz,
                c.qmuLRF_x, c.qmuLRF_y, c.qmuLRF_z};
            memcpy(dst, rec, sizeof(rec));
            dst += ISS_NFIELD;
        }
    }
    return static_cast<int64_t>(surf.size());
}

int32_t iss_host_species(iss_host *s, iss_species *dst) {
    GpuFSSW *g = s->obj->get_sampler();
    if (!g) return 0;
    const auto &sp = g->species();
    if (dst) memcpy(dst, sp.data(), sizeof(iss_species)*sp.size());
    return static_cast<int32_t>(sp.size());
}

const iss_hadron *iss_host_hadron_buffer(iss_host *s, const int64_t **event_offsets, int64_t *nev) {
    GpuFSSW *g = s->obj->get_sampler();
    if (!g) return nullptr;
    if (event_offsets) *event_offsets = g->event_offsets().data();
    if (nev) *nev = g->get_number_of_sampled_events();
    return reinterpret_cast<const iss_hadron *>(g->hadron_buffer());
}

int32_t iss_host_species_dN(iss_host *s, double *dst) {
    GpuFSSW *g = s->obj->get_sampler();
    if (!g) return 0;
    const auto &v = g->species_dN();
    if (dst) memcpy(dst, v.data(), sizeof(double)*v.size());
    return static_cast<int32_t>(v.size());
}

int iss_host_qa_block(iss_host *s, double *dst) {
    GpuFSSW *g = s->obj->get_sampler();
    if (!g || g->qa_block().empty()) return 1;
    memcpy(dst, g->qa_block().data(), sizeof(double)*g->qa_block().size());
    return 0;
}

int iss_host_spectra_table(iss_host *s, int32_t monval, double *dst, int32_t *npT, int32_t *nphi,
                           double *kernel_ms, double *evaluations) {
    GpuSpectra *g = s->obj->get_spectra();
    if (!g) return 1;
    if (npT) *npT = g->pT_tab_length();
    if (nphi) *nphi = g->phi_tab_length();
    if (kernel_ms) *kernel_ms = g->last_kernel_ms();
    if (evaluations) *evaluations = g->last_evaluations();
    const auto &pt = s->obj->get_particle_table();
    for (size_t n = 0; n < pt.size(); n++)
        if (pt[n].monval == monval) {
            const std::vector<double> &t = g->dN_pTdpTdphidy(static_cast<int>(n));
            if (t.empty()) return 1;
            if (dst) memcpy(dst, t.data(), sizeof(double)*t.size());
            return 0;
        }
    return 1;
}

int iss_host_write_samples(int format, const iss_hadron *hadrons, const int64_t *event_offsets,
                           int64_t nev, const char *table_path) {
    const iSS_Hadron *h = reinterpret_cast<const iSS_Hadron *>(hadrons);
    if (format == 0) {
        iss_writers::write_oscar("OSCAR.DAT", std::string(table_path ? table_path : "iSS_tables")
                                                  + "/OSCAR_header.txt", h, event_offsets, nev);
    } else if (format == 1) {
        iss_writers::write_gzip("particle_samples.gz", h, event_offsets, nev);
    } else if (format == 2) {
        iss_writers::write_binary("particle_samples.bin", h, event_offsets, nev);
    } else {
        return 1;
    }
    return 0;
}

}  // extern "C"
