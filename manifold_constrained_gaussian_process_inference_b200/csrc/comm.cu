// Multi-GPU collectives of the sampler: chains shard over the GPUs of one box with no collective on
// the hot path; the two exchange steps there are -- the pooled window statistics of the warm-up (one all-reduce per
// adaptation window) and the end-of-run all-gather of the retained (theta, sigma, lp) draws for R-hat / ESS -- run in-library
// on the sampler's stream through NCCL over NVLink.  libnccl is resolved at run time (dlopen of libnccl.so.2: the copy the host
// process already loaded, e.g. torch's, is reused), so the library keeps loading on hosts without NCCL; a communicator is
// either created here from a unique id the host distributes (magi_comm_init) or handed in by the host (magi_comm_attach).
#include <dlfcn.h>
#include <nccl.h>
#include <cstring>
#include "magi_internal.cuh"

namespace magi {

struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};

static NcclApi* nccl_api() {
    static NcclApi api;
    static bool tried = false;
    if (tried) return api.ok ? &api : nullptr;
    tried = true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) { api.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL); if (api.lib) break; }
    if (!api.lib) return nullptr;
#define MAGI_NCCL_SYM(field, sym) *(void**)(&api.field) = dlsym(api.lib, sym); if (!api.field) return nullptr;
    MAGI_NCCL_SYM(GetUniqueId, "ncclGetUniqueId") MAGI_NCCL_SYM(CommInitRank, "ncclCommInitRank") MAGI_NCCL_SYM(CommDestroy, "ncclCommDestroy")
    MAGI_NCCL_SYM(AllReduce, "ncclAllReduce") MAGI_NCCL_SYM(AllGather, "ncclAllGather") MAGI_NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef MAGI_NCCL_SYM
    api.ok = true;
    return &api;
}

static int nccl_error(NcclApi* a, ncclResult_t r, const char* what) {
    return set_error(MAGI_ERR_CUDA, std::string(what) + ": NCCL " + (a && a->GetErrorString ? a->GetErrorString(r) : "error"));
}

void comm_free(magi_handle* h) {
    if (h->nccl_comm && h->nccl_owned) { NcclApi* a = nccl_api(); if (a) a->CommDestroy((ncclComm_t)h->nccl_comm); }
    h->nccl_comm = nullptr; h->nccl_owned = false; h->nccl_world = 1; h->nccl_rank = 0;
}

// in-place sum over the ranks of n doubles on `st` (no-op for a single rank)
int comm_allreduce_sum(magi_handle* h, double* buf, size_t n, cudaStream_t st) {
    if (!h->nccl_comm || h->nccl_world <= 1) return MAGI_OK;
    NcclApi* a = nccl_api();
    if (!a) return set_error(MAGI_ERR_UNSUPPORTED, "libnccl.so.2 not found");
    ncclResult_t r = a->AllReduce(buf, buf, n, ncclDouble, ncclSum, (ncclComm_t)h->nccl_comm, st);
    return r == ncclSuccess ? MAGI_OK : nccl_error(a, r, "ncclAllReduce");
}

int comm_allgather(magi_handle* h, const double* send, double* recv, size_t n_per_rank, cudaStream_t st) {
    NcclApi* a = nccl_api();
    if (!a) return set_error(MAGI_ERR_UNSUPPORTED, "libnccl.so.2 not found");
    ncclResult_t r = a->AllGather(send, recv, n_per_rank, ncclDouble, (ncclComm_t)h->nccl_comm, st);
    return r == ncclSuccess ? MAGI_OK : nccl_error(a, r, "ncclAllGather");
}

}  // namespace magi

using namespace magi;

extern "C" int magi_nccl_unique_id(char* id /* MAGI_NCCL_ID_BYTES */) {
    if (!id) return set_error(MAGI_ERR_INVALID_ARGUMENT, "magi_nccl_unique_id: null argument");
    NcclApi* a = nccl_api();
    if (!a) return set_error(MAGI_ERR_UNSUPPORTED, "magi_nccl_unique_id: libnccl.so.2 not found");
    ncclUniqueId u;
    ncclResult_t r = a->GetUniqueId(&u);
    if (r != ncclSuccess) return nccl_error(a, r, "ncclGetUniqueId");
    static_assert(sizeof(u) == MAGI_NCCL_ID_BYTES, "ncclUniqueId size");
    memcpy(id, &u, sizeof u);
    return MAGI_OK;
}

extern "C" int magi_comm_init(magi_handle* h, const char* id, int rank, int world) {
    if (!h || !id || world < 1 || rank < 0 || rank >= world) return set_error(MAGI_ERR_INVALID_ARGUMENT, "magi_comm_init: bad argument");
    NcclApi* a = nccl_api();
    if (!a) return set_error(MAGI_ERR_UNSUPPORTED, "magi_comm_init: libnccl.so.2 not found");
    if (cudaSetDevice(h->device) != cudaSuccess) return set_error(MAGI_ERR_CUDA, "cudaSetDevice failed");
    comm_free(h);
    ncclUniqueId u;
    memcpy(&u, id, sizeof u);
    ncclComm_t c = nullptr;
    ncclResult_t r = a->CommInitRank(&c, world, u, rank);
    if (r != ncclSuccess) return nccl_error(a, r, "ncclCommInitRank");
    h->nccl_comm = c; h->nccl_owned = true; h->nccl_rank = rank; h->nccl_world = world;
    return MAGI_OK;
}

extern "C" int magi_comm_attach(magi_handle* h, void* nccl_comm, int rank, int world) {
    if (!h || world < 1 || rank < 0 || rank >= world) return set_error(MAGI_ERR_INVALID_ARGUMENT, "magi_comm_attach: bad argument");
    comm_free(h);
    h->nccl_comm = nccl_comm; h->nccl_owned = false; h->nccl_rank = rank; h->nccl_world = world;
    return MAGI_OK;
}

// The first collectives on a communicator set up its channels and staging buffers (measured on 2 B200s: the first 189 MB all-gather
// 16.7 ms, the second 1.4 ms, from the third on 0.26 ms = 725 GB/s per GPU): two rounds of an all-reduce and a 32 MB-per-rank
// all-gather before anything is timed
extern "C" int magi_comm_warmup(magi_handle* h, void* stream) {
    if (!h) return set_error(MAGI_ERR_INVALID_ARGUMENT, "magi_comm_warmup: null handle");
    if (!h->nccl_comm || h->nccl_world <= 1) return MAGI_OK;
    if (cudaSetDevice(h->device) != cudaSuccess) return set_error(MAGI_ERR_CUDA, "cudaSetDevice failed");
    cudaStream_t st = stream ? (cudaStream_t)stream : h->stream;
    double* buf = nullptr;
    const size_t n = (size_t)4 << 20;
    if (cudaMalloc(&buf, sizeof(double) * n * (size_t)h->nccl_world) != cudaSuccess) return set_error(MAGI_ERR_CUDA, "cudaMalloc failed");
    cudaMemsetAsync(buf, 0, sizeof(double) * n * (size_t)h->nccl_world, st);
    int rc = MAGI_OK;
    for (int round = 0; round < 3 && rc == MAGI_OK; ++round) {
        rc = comm_allreduce_sum(h, buf, 1 << 16, st);
        if (rc == MAGI_OK) rc = comm_allgather(h, buf + (size_t)h->nccl_rank * n, buf, n, st);
    }
    cudaError_t e = cudaStreamSynchronize(st);
    cudaFree(buf);
    if (rc != MAGI_OK) return rc;
    return e == cudaSuccess ? MAGI_OK : cuda_error(e, "magi_comm_warmup");
}
