// Multi-GPU plumbing for the z-sharded phase retrieval: one NCCL all-reduce (sum, float64) of
// the 2*(K*K+1) partial-pupil / squared-error sums per iteration.
// NCCL is bound at run time (dlopen of the libnccl the host process -- torch -- already uses),
// so the library has no link-time dependency on it and single-GPU use never touches it.
#include <dlfcn.h>

#include <cstring>
#include <mutex>

#include "common.cuh"

namespace {
// the slice of nccl.h this file needs (NCCL 2.x ABI)
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclFloat64 = 8 };
enum { ncclSum = 0 };

struct NcclApi {
    void* so = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi g_nccl;
std::mutex g_nccl_mu;
}  // namespace

struct pyotf_comm {
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1;
};

namespace pyotf {
int comm_allreduce_f64(pyotf_comm* c, double* buf, long long n, cudaStream_t st) {
    PYOTF_REQUIRE(c && c->comm && g_nccl.AllReduce, "comm: not initialised");
    ncclResult_t r = g_nccl.AllReduce(buf, buf, (size_t)n, ncclFloat64, ncclSum, c->comm, st);
    if (r != 0) { set_error(std::string("ncclAllReduce: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "error")); return 3; }
    return 0;
}
}  // namespace pyotf

using namespace pyotf;

extern "C" {

int pyotf_b200_comm_load(const char* libnccl_path) {
    std::lock_guard<std::mutex> lk(g_nccl_mu);
    if (g_nccl.so) return 0;
    void* so = dlopen(libnccl_path && libnccl_path[0] ? libnccl_path : "libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!so) { set_error(std::string("comm_load: ") + dlerror()); return 1; }
    NcclApi a;
    a.so = so;
    a.GetUniqueId = (decltype(a.GetUniqueId))dlsym(so, "ncclGetUniqueId");
    a.CommInitRank = (decltype(a.CommInitRank))dlsym(so, "ncclCommInitRank");
    a.CommDestroy = (decltype(a.CommDestroy))dlsym(so, "ncclCommDestroy");
    a.AllReduce = (decltype(a.AllReduce))dlsym(so, "ncclAllReduce");
    a.GetErrorString = (decltype(a.GetErrorString))dlsym(so, "ncclGetErrorString");
    if (!a.GetUniqueId || !a.CommInitRank || !a.CommDestroy || !a.AllReduce) {
        set_error("comm_load: the library does not export the NCCL entry points");
        dlclose(so);
        return 1;
    }
    g_nccl = a;
    return 0;
}

int pyotf_b200_comm_unique_id(unsigned char id_out[128]) {
    PYOTF_REQUIRE(g_nccl.GetUniqueId, "comm_unique_id: call pyotf_b200_comm_load first");
    ncclUniqueId id;
    ncclResult_t r = g_nccl.GetUniqueId(&id);
    if (r != 0) { set_error("ncclGetUniqueId failed"); return 3; }
    std::memcpy(id_out, id.internal, 128);
    return 0;
}

int pyotf_b200_comm_create(const unsigned char id_in[128], int rank, int world, pyotf_comm_t** out) {
    PYOTF_REQUIRE(g_nccl.CommInitRank, "comm_create: call pyotf_b200_comm_load first");
    PYOTF_REQUIRE(out && id_in && world >= 1 && rank >= 0 && rank < world, "comm_create: bad argument");
    ncclUniqueId id;
    std::memcpy(id.internal, id_in, 128);
    pyotf_comm* c = new pyotf_comm();
    c->rank = rank; c->world = world;
    ncclResult_t r = g_nccl.CommInitRank(&c->comm, world, id, rank);
    if (r != 0) {
        set_error(std::string("ncclCommInitRank: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "error"));
        delete c;
        return 3;
    }
    *out = c;
    return 0;
}

int pyotf_b200_comm_destroy(pyotf_comm_t* c) {
    if (!c) return 0;
    if (c->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm);
    delete c;
    return 0;
}

int pyotf_b200_comm_allreduce_f64(pyotf_comm_t* c, double* buf_dev, long long n, void* stream) {
    PYOTF_REQUIRE(buf_dev && n > 0, "comm_allreduce_f64: bad argument");
    return comm_allreduce_f64(c, buf_dev, n, (cudaStream_t)stream);
}

}  // extern "C"
