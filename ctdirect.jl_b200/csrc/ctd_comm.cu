// Multi-GPU norm gather of the batched path: one process per GPU, instances sharded over the
// ranks, and ONE exchange per batched evaluation -- the all-gather of the per-instance (objective,
// violation) pairs.  The collective is NCCL's (NVLink / NVSwitch); it is issued from here, on a stream owned
// by the handle, so that a Julia (or any non-torch) host needs nothing but the C ABI, and so that the
// per-evaluation host cost is one C call.  NCCL is loaded at run time (dlopen): the library has no link
// dependency on it and single-GPU users never touch it.
//
// Pipelining: the gather runs on its own high-priority stream behind an event.  Slot s has an event that
// fires when its gather has landed; ctd_comm_wait(h, s) makes the evaluation stream wait for it.  Waiting
// right after the gather gives the serial behaviour; waiting only before slot s is reused lets the gather
// of evaluation k overlap the kernel of evaluation k+1.
#include <dlfcn.h>

#include <cstdlib>
#include <cstring>
#include <mutex>

#include "ctd_model.hpp"

namespace ctd {

#define CTD_CUDA(call)                                                                         \
    do {                                                                                       \
        cudaError_t _e = (call);                                                               \
        if (_e != cudaSuccess)                                                                 \
            throw CtdError{(int)_e, std::string("CUDA error: ") + cudaGetErrorString(_e) + " at " #call}; \
    } while (0)

namespace {
// the few NCCL entry points that are used (their C ABI is stable across NCCL 2.x)
struct NcclUniqueId { char internal[kNcclIdBytes]; };
using NcclComm = void*;
constexpr int kNcclFloat64 = 8;   // ncclFloat64 / ncclDouble
struct NcclApi {
    void* lib = nullptr;
    int (*GetUniqueId)(NcclUniqueId*) = nullptr;
    int (*CommInitRank)(NcclComm*, int, NcclUniqueId, int) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, NcclComm, cudaStream_t) = nullptr;
    int (*CommDestroy)(NcclComm) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    int (*GetVersion)(int*) = nullptr;
};

NcclApi& nccl() {
    static NcclApi api;
    static std::once_flag once;
    static std::string err;
    std::call_once(once, [] {
        const char* env = getenv("CTD_NCCL_LIB");
        void* lib = nullptr;
        if (env && *env) lib = dlopen(env, RTLD_NOW | RTLD_GLOBAL);
        // a copy that the process has already loaded (e.g. the one bundled with torch) is preferred to a second one
        if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
        if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!lib) { err = std::string("cannot load libnccl.so.2 (set CTD_NCCL_LIB): ") + dlerror(); return; }
        api.lib = lib;
        api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(lib, "ncclGetUniqueId");
        api.CommInitRank = (decltype(api.CommInitRank))dlsym(lib, "ncclCommInitRank");
        api.AllGather = (decltype(api.AllGather))dlsym(lib, "ncclAllGather");
        api.CommDestroy = (decltype(api.CommDestroy))dlsym(lib, "ncclCommDestroy");
        api.GetErrorString = (decltype(api.GetErrorString))dlsym(lib, "ncclGetErrorString");
        api.GetVersion = (decltype(api.GetVersion))dlsym(lib, "ncclGetVersion");
        if (!api.GetUniqueId || !api.CommInitRank || !api.AllGather || !api.CommDestroy || !api.GetErrorString) {
            err = "libnccl.so.2 lacks an expected symbol";
            api.lib = nullptr;
        }
    });
    if (!api.lib) throw CtdError{-7, err.empty() ? std::string("NCCL is not available") : err};
    return api;
}

void nccl_check(int r, const char* what) {
    if (r != 0) throw CtdError{1000 + r, std::string("NCCL error in ") + what + ": " + nccl().GetErrorString(r)};
}
}  // namespace

void nccl_unique_id(char* id) {
    NcclUniqueId u;
    std::memset(&u, 0, sizeof(u));
    nccl_check(nccl().GetUniqueId(&u), "ncclGetUniqueId");
    std::memcpy(id, u.internal, kNcclIdBytes);
}

int nccl_version() {
    int v = 0;
    if (nccl().GetVersion) nccl().GetVersion(&v);
    return v;
}

void Model::comm_init(const char* id, int rank, int world) {
    if (comm_) throw CtdError{-2, "ctd_comm_init: this handle already has a communicator"};
    if (!id || world < 1 || rank < 0 || rank >= world) throw CtdError{-2, "ctd_comm_init: bad id / rank / world"};
    ensure_device();
    CTD_CUDA(cudaSetDevice(device_));
    NcclUniqueId u;
    std::memcpy(u.internal, id, kNcclIdBytes);
    NcclComm c = nullptr;
    nccl_check(nccl().CommInitRank(&c, world, u, rank), "ncclCommInitRank");
    comm_ = c;
    comm_rank_ = rank;
    comm_world_ = world;
    int lo = 0, hi = 0;
    CTD_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));   // hi = numerically lowest = highest priority
    CTD_CUDA(cudaStreamCreateWithPriority(&comm_stream_, cudaStreamNonBlocking, hi));
    CTD_CUDA(cudaEventCreateWithFlags(&ev_issue_, cudaEventDisableTiming));
    for (int s = 0; s < kCommSlots; ++s) {
        CTD_CUDA(cudaEventCreateWithFlags(&ev_gathered_[s], cudaEventDisableTiming));
        slot_used_[s] = false;
    }
}

void Model::comm_destroy() {
    if (!comm_) return;
    cudaSetDevice(device_);
    if (comm_stream_) cudaStreamSynchronize(comm_stream_);
    try { nccl().CommDestroy(comm_); } catch (...) {}
    comm_ = nullptr;
    if (ev_issue_) cudaEventDestroy(ev_issue_);
    for (int s = 0; s < kCommSlots; ++s)
        if (ev_gathered_[s]) cudaEventDestroy(ev_gathered_[s]);
    if (comm_stream_) cudaStreamDestroy(comm_stream_);
    comm_stream_ = nullptr;
}

void Model::gather_norms_dev(const double* local, double* gathered, int slot) {
    if (!comm_) throw CtdError{-2, "ctd_gather_norms_dev: call ctd_comm_init first"};
    if (!local || !gathered) throw CtdError{-2, "ctd_gather_norms_dev: NULL buffer"};
    if (slot < 0 || slot >= kCommSlots) throw CtdError{-2, "ctd_gather_norms_dev: bad slot"};
    CTD_CUDA(cudaSetDevice(device_));
    // the collective starts when everything issued so far on the evaluation stream has finished
    CTD_CUDA(cudaEventRecord(ev_issue_, stream_));
    CTD_CUDA(cudaStreamWaitEvent(comm_stream_, ev_issue_, 0));
    nccl_check(nccl().AllGather(local, gathered, (size_t)2 * batch_, kNcclFloat64, comm_, comm_stream_), "ncclAllGather");
    CTD_CUDA(cudaEventRecord(ev_gathered_[slot], comm_stream_));
    slot_used_[slot] = true;
}

void Model::comm_wait(int slot) {
    if (!comm_) throw CtdError{-2, "ctd_comm_wait: call ctd_comm_init first"};
    if (slot < 0 || slot >= kCommSlots) throw CtdError{-2, "ctd_comm_wait: bad slot"};
    if (slot_used_[slot]) CTD_CUDA(cudaStreamWaitEvent(stream_, ev_gathered_[slot], 0));
}

}  // namespace ctd
