// engine_pool.h -- process-wide caches shared by the host drivers (gpu_fssw.cpp, gpu_spectra.cpp):
// CUDA handles by device and parsed coefficient tables.  MUSIC/JETSCAPE-style hosts call
// iSS::generate_samples() once per hydro event; the handles and tables outlive the samplers.
#ifndef ISS_B200_ENGINE_POOL_H_
#define ISS_B200_ENGINE_POOL_H_

#include <cstdint>
#include <string>
#include <vector>

#include "../../include/iss_cuda.h"

namespace iss_pool {

// LOCAL_RANK (torchrun) or ISS_CUDA_DEVICE, else 0
int default_device();
// an idle handle of the device or a new one; exits when no CUDA device is usable (no CPU fallback)
iss_handle *acquire_handle(int device);
void release_handle(int device, iss_handle *h);
// pinned host memory, pooled (cudaHostAlloc costs ~0.2 ms per MB): a block of at least `bytes`
struct PinnedBlock {
    void *ptr = nullptr;
    int64_t bytes = 0;
};
PinnedBlock pinned_acquire(iss_handle *h, int64_t bytes);
void pinned_release(iss_handle *h, PinnedBlock b);
// numbers of a whitespace separated text file after `skip_lines` header lines, parsed once
const std::vector<double> &cached_numbers(const std::string &file, int skip_lines);

}  // namespace iss_pool

#endif  // ISS_B200_ENGINE_POOL_H_
