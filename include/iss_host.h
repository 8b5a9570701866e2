/*
 * iss_host.h -- C binding of the drop-in C++ facade `class iSS` (iss_b200/host/iSS.h, same
 * public API as reference src/iSS.h:16-102) for callers that cannot include a C++ header:
 * the Python tests, bench.py (ctypes) and C hosts.  One function per public member of the
 * reference class; each comment names the member it forwards to.  Fatal errors keep the
 * reference's convention (message + exit), see SURVEY.md section 8(b).
 */
#ifndef ISS_HOST_H_
#define ISS_HOST_H_

#include <stdint.h>

#include "iss_cuda.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct iss_host iss_host;

/* iSS::iSS(path, table_path, particle_table_path, inputfile, surface_filename)  (iSS.h:43-46) */
ISS_API iss_host *iss_host_create(const char *path, const char *table_path,
                                  const char *particle_table_path, const char *inputfile,
                                  const char *surface_filename);
ISS_API void iss_host_destroy(iss_host *s);                          /* iSS::~iSS           */
/* paraRdr_ptr->setVal / getVal / phraseOneLine("key=value")  (iSS.h:49) */
ISS_API void iss_host_set_param(iss_host *s, const char *name, double value);
ISS_API double iss_host_get_param(iss_host *s, const char *name, double default_value);
ISS_API void iss_host_parse_param(iss_host *s, const char *key_equals_value);
ISS_API void iss_host_set_random_seed(iss_host *s, int seed);        /* iSS::set_random_seed */
ISS_API int iss_host_read_in_FO_surface(iss_host *s);                /* iSS::read_in_FO_surface */
ISS_API int iss_host_generate_samples(iss_host *s);                  /* iSS::generate_samples  */
ISS_API int iss_host_shell(iss_host *s);                             /* iSS::shell            */
ISS_API void iss_host_perform_checks(iss_host *s);                   /* iSS::perform_checks   */
ISS_API int iss_host_get_number_of_sampled_events(iss_host *s);      /* iSS.h:61 */
ISS_API int iss_host_get_number_of_particles(iss_host *s, int iev);  /* iSS.h:69 */
/* iSS::get_hadron_list_iev: pointer to n contiguous 40-byte records, owned by the sampler */
ISS_API const iss_hadron *iss_host_get_hadron_list_iev(iss_host *s, int iev, int64_t *n);
ISS_API void iss_host_clear(iss_host *s);                            /* iSS::clear */

/* ---- additions of the B200 engine ---------------------------------------------------- */
/* builds the device sampler (uploads surface, species, tables) without sampling */
ISS_API int iss_host_prepare_sampler(iss_host *s);
/* the CUDA handle of the prepared sampler, for direct use of include/iss_cuda.h */
ISS_API iss_handle *iss_host_cuda_handle(iss_host *s);
/* local-rest-frame surface held by the facade: ncell, and (if dst != NULL) ncell x 28 floats
 * in ISS_F_* order */
ISS_API int64_t iss_host_lrf_surface(iss_host *s, float *dst);
/* chosen species in sampling order (if dst != NULL, filled with nspecies records) */
ISS_API int32_t iss_host_species(iss_host *s, iss_species *dst);
/* contiguous pinned buffer of all events + event offsets [nev+1] after generate_samples */
ISS_API const iss_hadron *iss_host_hadron_buffer(iss_host *s, const int64_t **event_offsets,
                                                 int64_t *nev);
/* per-species dN (sum over cells) of the last yield computation, nspecies doubles */
ISS_API int32_t iss_host_species_dN(iss_host *s, double *dst);
/* QA block (layout: iss_cuda.h), iss_cuda_qa_size() doubles */
ISS_API int iss_host_qa_block(iss_host *s, double *dst);

/* smooth spectra (MC_sampling = 0, calculate_vn = 1): the dN/(pT dpT dphi dy) table of one species
 * after generate_samples, [npT][nphi] doubles (EmissionFunctionArray::dN_pTdpTdphidy,
 * emissionfunction.h:64).  dst may be NULL to query the sizes.  Returns 0, or 1 when the species
 * was not calculated.  kernel_ms / evaluations (may be NULL): device time and integrand
 * evaluations of the run. */
ISS_API int iss_host_spectra_table(iss_host *s, int32_t monval, double *dst, int32_t *npT,
                                   int32_t *nphi, double *kernel_ms, double *evaluations);

/* the reference's sample-file writers (FSSW::combine_samples_to_OSCAR / _gzip_file / _binary_file,
 * FSSW.cpp:365-561) on a caller-supplied hadron list; files go to the current directory like the
 * reference's.  format: 0 OSCAR.DAT, 1 particle_samples.gz, 2 particle_samples.bin             */
ISS_API int iss_host_write_samples(int format, const iss_hadron *hadrons,
                                   const int64_t *event_offsets, int64_t nev,
                                   const char *table_path);

#ifdef __cplusplus
}
#endif
#endif  /* ISS_HOST_H_ */
