// Multi-GPU plumbing of the C ABI: the ONE collective of the path -- the sum all-reduce of the packed float64
// Baum-Welch count buffer before the M-step (the reference has no multi-process code at all, the
// statistics it sums are those of DiscreteHMM.do_pass, discrete_hmm.py:117-159).
//
// NCCL is bound at run time (dlopen of libnccl.so.2, preferring the copy already loaded into the process, e.g. the
// one PyTorch ships) so libkhmm.so has no link-time dependency on it and single-GPU users never touch it.  The
// prototypes below restate the stable NCCL 2.x C API (nccl.h: ncclGetUniqueId / ncclCommInitRank / ncclAllReduce /
// ncclCommDestroy / ncclGetErrorString / ncclGetVersion).
#include <dlfcn.h>
#include <stdlib.h>

#include "common.cuh"

namespace khmm {

struct NcclUniqueId { char internal[128]; };
typedef void *NcclComm;
typedef int (*PFN_GetUniqueId)(NcclUniqueId *);
typedef int (*PFN_CommInitRank)(NcclComm *, int, NcclUniqueId, int);
typedef int (*PFN_AllReduce)(const void *, void *, size_t, int, int, NcclComm, cudaStream_t);
typedef int (*PFN_CommDestroy)(NcclComm);
typedef const char *(*PFN_GetErrorString)(int);
typedef int (*PFN_GetVersion)(int *);
typedef int (*PFN_CommCount)(NcclComm, int *);

constexpr int NCCL_FLOAT64 = 8, NCCL_SUM = 0;

struct NcclApi {
    void *handle = nullptr;
    PFN_GetUniqueId get_unique_id = nullptr;
    PFN_CommInitRank comm_init_rank = nullptr;
    PFN_AllReduce all_reduce = nullptr;
    PFN_CommDestroy comm_destroy = nullptr;
    PFN_GetErrorString get_error_string = nullptr;
    PFN_GetVersion get_version = nullptr;
    PFN_CommCount comm_count = nullptr;
};

static NcclApi g_nccl;

static int nccl_load() {
    if (g_nccl.handle) return KHMM_OK;
    const char *env = getenv("KHMM_NCCL_LIB");
    void *h = nullptr;
    if (env && *env) h = dlopen(env, RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);   // the copy the process already holds (PyTorch's)
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) {
        set_error("khmm_comm: libnccl.so.2 not found (%s); set KHMM_NCCL_LIB to its path", dlerror());
        return KHMM_ERR_UNSUPPORTED;
    }
    NcclApi a;
    a.handle = h;
    a.get_unique_id = (PFN_GetUniqueId)dlsym(h, "ncclGetUniqueId");
    a.comm_init_rank = (PFN_CommInitRank)dlsym(h, "ncclCommInitRank");
    a.all_reduce = (PFN_AllReduce)dlsym(h, "ncclAllReduce");
    a.comm_destroy = (PFN_CommDestroy)dlsym(h, "ncclCommDestroy");
    a.get_error_string = (PFN_GetErrorString)dlsym(h, "ncclGetErrorString");
    a.get_version = (PFN_GetVersion)dlsym(h, "ncclGetVersion");
    a.comm_count = (PFN_CommCount)dlsym(h, "ncclCommCount");
    if (!a.get_unique_id || !a.comm_init_rank || !a.all_reduce || !a.comm_destroy || !a.get_error_string) {
        set_error("khmm_comm: the NCCL library lacks a required symbol");
        return KHMM_ERR_UNSUPPORTED;
    }
    g_nccl = a;
    return KHMM_OK;
}

static int nccl_fail(int rc, const char *what) {
    set_error("NCCL error %d (%s) at %s", rc, g_nccl.get_error_string ? g_nccl.get_error_string(rc) : "?", what);
    return KHMM_ERR_CUDA;
}

struct Comm {
    NcclComm nccl;
    int rank, world, device;
};

}  // namespace khmm

using namespace khmm;

extern "C" {

int khmm_comm_nccl_version(int *version_host) {
    KHMM_REQUIRE(version_host, "comm_nccl_version: NULL argument");
    int rc = nccl_load();
    if (rc) return rc;
    *version_host = 0;
    if (g_nccl.get_version) g_nccl.get_version(version_host);
    return KHMM_OK;
}

int khmm_comm_unique_id(void *id128_host) {
    KHMM_REQUIRE(id128_host, "comm_unique_id: NULL argument");
    int rc = nccl_load();
    if (rc) return rc;
    NcclUniqueId id;
    int n = g_nccl.get_unique_id(&id);
    if (n) return nccl_fail(n, "ncclGetUniqueId");
    memcpy(id128_host, id.internal, KHMM_COMM_ID_BYTES);
    return KHMM_OK;
}

int khmm_comm_init_rank(const void *id128_host, int rank, int world, int device, void **comm_out) {
    KHMM_REQUIRE(id128_host && comm_out, "comm_init_rank: NULL argument");
    KHMM_REQUIRE(world >= 1 && rank >= 0 && rank < world, "comm_init_rank: rank %d outside [0, %d)", rank, world);
    int rc = nccl_load();
    if (rc) return rc;
    KHMM_CUDA_CHECK(cudaSetDevice(device));
    NcclUniqueId id;
    memcpy(id.internal, id128_host, KHMM_COMM_ID_BYTES);
    NcclComm c = nullptr;
    int n = g_nccl.comm_init_rank(&c, world, id, rank);
    if (n) return nccl_fail(n, "ncclCommInitRank");
    Comm *comm = new Comm{c, rank, world, device};
    *comm_out = comm;
    return KHMM_OK;
}

int khmm_comm_info(void *comm, int *rank_host, int *world_host, int *nccl_world_host) {
    KHMM_REQUIRE(comm, "comm_info: NULL communicator");
    Comm *c = (Comm *)comm;
    if (rank_host) *rank_host = c->rank;
    if (world_host) *world_host = c->world;
    if (nccl_world_host) {
        *nccl_world_host = -1;
        if (g_nccl.comm_count) {
            int n = g_nccl.comm_count(c->nccl, nccl_world_host);
            if (n) return nccl_fail(n, "ncclCommCount");
        }
    }
    return KHMM_OK;
}

int khmm_allreduce_counts(void *comm, double *counts, size_t n, khmm_stream_t stream) {
    KHMM_REQUIRE(comm && counts, "allreduce_counts: NULL argument");
    Comm *c = (Comm *)comm;
    if (c->world == 1 || n == 0) return KHMM_OK;
    int rc = g_nccl.all_reduce(counts, counts, n, NCCL_FLOAT64, NCCL_SUM, c->nccl, (cudaStream_t)stream);
    if (rc) return nccl_fail(rc, "ncclAllReduce");
    return KHMM_OK;
}

int khmm_comm_destroy(void *comm) {
    if (!comm) return KHMM_OK;
    Comm *c = (Comm *)comm;
    int rc = g_nccl.comm_destroy ? g_nccl.comm_destroy(c->nccl) : 0;
    delete c;
    if (rc) return nccl_fail(rc, "ncclCommDestroy");
    return KHMM_OK;
}

}  // extern "C"
