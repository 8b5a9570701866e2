// collective.cu -- the collectives of the hot path behind the C ABI: the QA block
// (iSS::perform_checks, reference src/iSS.cpp:59-83, 296-363: every entry is a plain sum over
// hadrons / events) summed over the ranks of a one-process-per-GPU job, and, in surface-chunk
// mode, the all-gather of the ranks' tile sums between the local and the global part of the yields.  NCCL is bound at run
// time (dlopen of libnccl.so.2), so single-GPU hosts need no NCCL at all; a null communicator
// means "local": the block is left as it is.
#include <dlfcn.h>

#include "iss_internal.cuh"

namespace {

// the slice of NCCL's C API used here (nccl.h of NCCL 2.x; the ABI of these five has been stable
// since 2.0)
struct NcclUniqueId { char internal[128]; };
typedef void *NcclComm;
typedef int (*GetUniqueIdFn)(NcclUniqueId *);
typedef int (*CommInitRankFn)(NcclComm *, int, NcclUniqueId, int);
typedef int (*AllReduceFn)(const void *, void *, size_t, int, int, NcclComm, cudaStream_t);
typedef int (*AllGatherFn)(const void *, void *, size_t, int, NcclComm, cudaStream_t);
typedef int (*CommDestroyFn)(NcclComm);
typedef const char *(*GetErrorStringFn)(int);
constexpr int NCCL_FLOAT64 = 8, NCCL_SUM = 0;

struct NcclApi {
    void *lib = nullptr;
    GetUniqueIdFn get_unique_id = nullptr;
    CommInitRankFn comm_init_rank = nullptr;
    AllReduceFn all_reduce = nullptr;
    AllGatherFn all_gather = nullptr;
    CommDestroyFn comm_destroy = nullptr;
    GetErrorStringFn error_string = nullptr;
    std::string err;
};

NcclApi &nccl() {
    static NcclApi api;
    if (api.lib || !api.err.empty()) return api;
    // a process that already carries NCCL (e.g. through torch) gets that copy: same soname
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *n : names) {
        api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (api.lib) break;
    }
    if (!api.lib) {
        api.err = std::string("libnccl.so.2 not found: ") + dlerror();
        return api;
    }
    api.get_unique_id = reinterpret_cast<GetUniqueIdFn>(dlsym(api.lib, "ncclGetUniqueId"));
    api.comm_init_rank = reinterpret_cast<CommInitRankFn>(dlsym(api.lib, "ncclCommInitRank"));
    api.all_reduce = reinterpret_cast<AllReduceFn>(dlsym(api.lib, "ncclAllReduce"));
    api.all_gather = reinterpret_cast<AllGatherFn>(dlsym(api.lib, "ncclAllGather"));
    api.comm_destroy = reinterpret_cast<CommDestroyFn>(dlsym(api.lib, "ncclCommDestroy"));
    api.error_string = reinterpret_cast<GetErrorStringFn>(dlsym(api.lib, "ncclGetErrorString"));
    if (!api.get_unique_id || !api.comm_init_rank || !api.all_reduce || !api.all_gather || !api.comm_destroy) {
        api.err = "libnccl.so.2 lacks ncclGetUniqueId / ncclCommInitRank / ncclAllReduce / ncclAllGather / "
                  "ncclCommDestroy";
        api.lib = nullptr;
    }
    return api;
}

std::string nccl_error(NcclApi &api, const char *what, int rc) {
    std::string s = std::string(what) + " failed";
    if (api.error_string) s += std::string(": ") + api.error_string(rc);
    return s;
}

}  // namespace

extern "C" {

int iss_cuda_nccl_unique_id(void *id128) {
    if (!id128) return ISS_ERR_ARG;
    NcclApi &api = nccl();
    if (!api.lib) return ISS_ERR_STATE;
    NcclUniqueId id;
    if (api.get_unique_id(&id) != 0) return ISS_ERR_CUDA;
    memcpy(id128, id.internal, sizeof(id.internal));
    return ISS_OK;
}

int iss_cuda_nccl_init(iss_handle *h, const void *id128, int32_t rank, int32_t nranks) {
    if (!h || !id128 || nranks <= 0 || rank < 0 || rank >= nranks) return ISS_ERR_ARG;
    NcclApi &api = nccl();
    if (!api.lib) ISS_FAIL(h, ISS_ERR_STATE, api.err);
    cudaSetDevice(h->device);
    if (h->nccl_comm) {
        api.comm_destroy(h->nccl_comm);
        h->nccl_comm = nullptr;
    }
    NcclUniqueId id;
    memcpy(id.internal, id128, sizeof(id.internal));
    NcclComm comm = nullptr;
    const int rc = api.comm_init_rank(&comm, nranks, id, rank);
    if (rc != 0) ISS_FAIL(h, ISS_ERR_CUDA, nccl_error(api, "ncclCommInitRank", rc));
    h->nccl_comm = comm;
    h->nccl_rank = rank;
    h->nccl_nranks = nranks;
    return ISS_OK;
}

int iss_cuda_nccl_finalize(iss_handle *h) {
    if (!h) return ISS_ERR_ARG;
    if (h->nccl_comm) {
        cudaSetDevice(h->device);
        cudaStreamSynchronize(h->stream);
        nccl().comm_destroy(h->nccl_comm);
        h->nccl_comm = nullptr;
        h->nccl_nranks = 0;
    }
    return ISS_OK;
}

int iss_cuda_histograms_allreduce(iss_handle *h, void *nccl_comm) {
    if (!h) return ISS_ERR_ARG;
    if (!h->d_qa) ISS_FAIL(h, ISS_ERR_STATE, "no QA block (call iss_cuda_histograms)");
    NcclComm comm = nccl_comm ? nccl_comm : h->nccl_comm;
    if (!comm) return ISS_OK;       // one rank: the local block is the global one
    NcclApi &api = nccl();
    if (!api.lib) ISS_FAIL(h, ISS_ERR_STATE, api.err);
    cudaSetDevice(h->device);
    const int rc = api.all_reduce(h->d_qa, h->d_qa, static_cast<size_t>(iss_cuda_qa_size()),
                                  NCCL_FLOAT64, NCCL_SUM, comm, h->stream);
    if (rc != 0) ISS_FAIL(h, ISS_ERR_CUDA, nccl_error(api, "ncclAllReduce", rc));
    return ISS_OK;
}

int iss_cuda_chunk_yields_allgather(iss_handle *h, const int64_t *rank_ntile, int32_t nranks,
                                    void *nccl_comm, double *dN_species_host) {
    if (!h || !rank_ntile || nranks <= 0) return ISS_ERR_ARG;
    if (!h->chunk) ISS_FAIL(h, ISS_ERR_STATE, "iss_cuda_set_surface_chunk must run first");
    NcclComm comm = nccl_comm ? nccl_comm : h->nccl_comm;
    if (nranks > 1 && !comm)
        ISS_FAIL(h, ISS_ERR_STATE, "no communicator: pass an ncclComm_t or call iss_cuda_nccl_init first");
    cudaSetDevice(h->device);
    // equal blocks of [nspecies][width] doubles per rank: a rank's [nspecies][ntile] table is the
    // head of its block (the send buffer is the tile-sum table itself, allocated with that size)
    int64_t width = 0;
    for (int r = 0; r < nranks; r++) {
        if (rank_ntile[r] < 0) return ISS_ERR_ARG;
        width = std::max(width, rank_ntile[r]);
    }
    const int64_t ns = h->nspecies;
    const size_t block = static_cast<size_t>(ns)*static_cast<size_t>(width);
    if (sizeof(double)*block > h->tilesum_bytes || !h->d_tilesum) {
        ISS_ENSURE(h, h->d_tilesum, h->tilesum_bytes, sizeof(double)*block);
        ISS_CUDA_TRY(h, cudaMemsetAsync(h->d_tilesum, 0, sizeof(double)*block, h->stream));
    }
    int rc = iss::run_yields_local(h);
    if (rc) return rc;
    if (h->ntile > width) ISS_FAIL(h, ISS_ERR_ARG, "rank_ntile does not list this handle's chunk");
    std::vector<const double *> blocks(nranks);
    if (nranks == 1) {
        blocks[0] = h->d_tilesum;
    } else {
        NcclApi &api = nccl();
        if (!api.lib) ISS_FAIL(h, ISS_ERR_STATE, api.err);
        ISS_ENSURE(h, h->d_tilesum_all, h->tilesum_all_bytes, sizeof(double)*block*nranks);
        const int nrc = api.all_gather(h->d_tilesum, h->d_tilesum_all, block, NCCL_FLOAT64, comm, h->stream);
        if (nrc != 0) ISS_FAIL(h, ISS_ERR_CUDA, nccl_error(api, "ncclAllGather", nrc));
        for (int r = 0; r < nranks; r++) blocks[r] = h->d_tilesum_all + block*r;
    }
    return iss::chunk_combine_tile_sums(h, blocks.data(), rank_ntile, nranks, 1, dN_species_host);
}

}  // extern "C"
