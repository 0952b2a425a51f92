// Multi-GPU norm gather of the batched path: one process per GPU, instances sharded over the
// ranks, and ONE exchange per batched evaluation -- the all-gather of the per-instance (objective,
// violation) pairs.  The collective is NCCL's (NVLink / NVSwitch); it is issued from here, on a stream owned
// by the handle, so that a Julia (or any non-torch) host needs nothing but the C ABI, and so that the
// per-evaluation host cost is one C call.  NCCL is loaded at run time (dlopen): the library has no link
// dependency on it and single-GPU users never touch it.
//
// Fused gather over peer memory (ctd_peer_init / ctd_eval_gather_dev / ctd_peer_wait): NCCL's all-gather of 64 KiB is
// 20-27 us of pure latency at 8 GPUs, against a 45 us evaluation kernel.  With peer buffers the evaluation kernel
// itself stores every instance's (objective, violation) into the gather buffer of EVERY rank as soon as the instance
// is reduced (posted NVLink peer stores from the reduction epilogue, ctd_kernels.cuh): the transfer overlaps the rest
// of the evaluation, instance by instance.  The values travel as self-validating 16-byte records keyed by the launch,
// so no fence, flag or counter is on the producer's path; the consumer side is a one-CTA kernel that polls the
// slot's records until they all validate and copies them out.  The buffers are exchanged once, as CUDA IPC handles
// over the NCCL communicator.
//
// Pipelining: the gather runs on its own high-priority stream behind an event.  Slot s has an event that
// fires when its gather has landed; ctd_comm_wait(h, s) makes the evaluation stream wait for it.  Waiting
// right after the gather gives the serial behaviour; waiting only before slot s is reused lets the gather
// of evaluation k overlap the kernel of evaluation k+1.
#include <dlfcn.h>

#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

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

// ---- fused gather over peer memory ----------------------------------------------------------------------------
// Layout of every rank's buffer, in doubles (all ranks have the same batch):
//   records  [slot][src rank][2][batch] x 2   self-validating {bits, bits ^ key}, written by the rank `src`
//   results  [slot][src rank][2][batch]       what ctd_peer_wait_kernel copies out for the consumers (local use)
//   acks     [slot][src rank]                 u64: rank `src` has copied round e of the slot out (written by `src`)
size_t Model::peer_records() const { return (size_t)kCommSlots * comm_world_ * 2 * batch_ * 2; }
size_t Model::peer_results() const { return (size_t)kCommSlots * comm_world_ * 2 * batch_; }

static unsigned long long peer_key_of(unsigned long long epoch, int slot) {   // splitmix64 finaliser; never 0
    unsigned long long e = epoch * 4ull + (unsigned long long)slot + 0x9E3779B97F4A7C15ull;
    e = (e ^ (e >> 30)) * 0xBF58476D1CE4E5B9ull;
    e = (e ^ (e >> 27)) * 0x94D049BB133111EBull;
    e ^= e >> 31;
    return e | 1ull;
}

void Model::peer_init() {
    if (!comm_) throw CtdError{-2, "ctd_peer_init: call ctd_comm_init first"};
    if (peer_local_) return;
    if (reg_.launcher.cfg(F_CONS | F_OBJ | F_VIOL).threads == 0)
        throw CtdError{-6, "ctd_peer_init: the fused gather needs the violation norm reduced inside the evaluation kernel "
                           "(light elements); use ctd_gather_norms_dev for this element"};
    CTD_CUDA(cudaSetDevice(device_));
    const int W = comm_world_;
    const size_t n = peer_records() + peer_results() + (size_t)kCommSlots * W;
    CTD_CUDA(cudaMalloc(&peer_local_, sizeof(double) * n));
    CTD_CUDA(cudaMemset(peer_local_, 0, sizeof(double) * n));
    CTD_CUDA(cudaMalloc(&d_peer_gate_, sizeof(unsigned long long) * kCommSlots));
    CTD_CUDA(cudaMemset(d_peer_gate_, 0, sizeof(unsigned long long) * kCommSlots));
    CTD_CUDA(cudaMalloc(&d_peer_err_, sizeof(int)));
    CTD_CUDA(cudaMemset(d_peer_err_, 0, sizeof(int)));
    // exchange the IPC handles of the buffers over the communicator
    cudaIpcMemHandle_t mine;
    CTD_CUDA(cudaIpcGetMemHandle(&mine, peer_local_));
    char *d_send = nullptr, *d_recv = nullptr;
    CTD_CUDA(cudaMalloc(&d_send, sizeof(mine)));
    CTD_CUDA(cudaMalloc(&d_recv, sizeof(mine) * W));
    CTD_CUDA(cudaMemcpy(d_send, &mine, sizeof(mine), cudaMemcpyHostToDevice));
    nccl_check(nccl().AllGather(d_send, d_recv, sizeof(mine), /*ncclChar*/ 0, comm_, comm_stream_), "ncclAllGather (IPC handles)");
    CTD_CUDA(cudaStreamSynchronize(comm_stream_));
    std::vector<cudaIpcMemHandle_t> all(W);
    CTD_CUDA(cudaMemcpy(all.data(), d_recv, sizeof(mine) * W, cudaMemcpyDeviceToHost));
    cudaFree(d_send);
    cudaFree(d_recv);
    peer_base_.assign(W, nullptr);
    for (int r = 0; r < W; ++r) {
        if (r == comm_rank_) { peer_base_[r] = peer_local_; continue; }
        void* ptr = nullptr;
        const cudaError_t e = cudaIpcOpenMemHandle(&ptr, all[r], cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess)
            throw CtdError{(int)e, std::string("ctd_peer_init: cannot map the buffer of rank ") + std::to_string(r) + ": " +
                                       cudaGetErrorString(e) + " (peer access over NVLink is required)"};
        peer_base_[r] = static_cast<double*>(ptr);
    }
    CTD_CUDA(cudaMalloc(&d_peer_base_, sizeof(double*) * W));
    CTD_CUDA(cudaMemcpy(d_peer_base_, peer_base_.data(), sizeof(double*) * W, cudaMemcpyHostToDevice));
    // every rank's buffer is zeroed and mapped before anybody pushes into it (a small in-place all-gather as barrier)
    int* d_bar = nullptr;
    CTD_CUDA(cudaMalloc(&d_bar, sizeof(int) * W));
    CTD_CUDA(cudaMemset(d_bar, 0, sizeof(int) * W));
    nccl_check(nccl().AllGather(d_bar + comm_rank_, d_bar, sizeof(int), /*ncclChar*/ 0, comm_, comm_stream_), "ncclAllGather (barrier)");
    CTD_CUDA(cudaStreamSynchronize(comm_stream_));
    cudaFree(d_bar);
    for (int s = 0; s < kCommSlots; ++s) peer_epoch_[s] = 0;
}

void Model::peer_destroy() {
    if (!peer_local_) return;
    cudaSetDevice(device_);
    cudaDeviceSynchronize();
    for (int r = 0; r < (int)peer_base_.size(); ++r)
        if (r != comm_rank_ && peer_base_[r]) cudaIpcCloseMemHandle(peer_base_[r]);
    cudaFree(peer_local_);
    cudaFree(d_peer_gate_);
    cudaFree(d_peer_err_);
    cudaFree(d_peer_base_);
    peer_local_ = nullptr;
    peer_base_.clear();
}

void Model::eval_gather_dev(int flags, const double* x, const double* y, double sigma, double* grad, double* c,
                            double* jvals, double* hvals, int slot) {
    if (!peer_local_) throw CtdError{-2, "ctd_eval_gather_dev: call ctd_peer_init first"};
    if (slot < 0 || slot >= kCommSlots) throw CtdError{-2, "ctd_eval_gather_dev: bad slot"};
    if (!(flags & F_CONS) || !(flags & F_OBJ)) throw CtdError{-2, "ctd_eval_gather_dev: CTD_CONS and CTD_OBJ must be among the outputs"};
    ensure_staging();
    ensure_con_bounds();
    PeerPush pp;
    pp.slot = slot;
    pp.epoch = ++peer_epoch_[slot];
    peer_push_ = &pp;
    try {
        eval_dev(flags, x, y, sigma, d_obj_, grad, c, jvals, hvals, d_viol_);
    } catch (...) {
        peer_push_ = nullptr;
        throw;
    }
    peer_push_ = nullptr;
}

void Model::peer_args(EvalArgs& a) const {
    const int s = peer_push_->slot;
    a.peer_base = d_peer_base_;
    a.peer_world = comm_world_;
    a.peer_rec = ((long long)s * comm_world_ + comm_rank_) * 2 * batch_ * 2;
    a.peer_epoch = peer_push_->epoch;
    a.peer_key = peer_key_of(peer_push_->epoch, s);
    a.peer_ack_local = reinterpret_cast<const unsigned long long*>(peer_local_ + peer_records() + peer_results()) + (size_t)s * comm_world_;
    a.peer_gate = d_peer_gate_ + s;
}

// The consumer side, one CTA: every thread polls its share of the slot's records until they validate for this round and
// copies the values into the slot's result rows; then the rank acknowledges the round to every rank (the slot may be
// pushed again).  Launched with the programmatic-stream-serialization attribute, so it is already resident and
// polling while the evaluation that pushes this rank's own records runs its last wave (that kernel never waits for
// anything this one does); it ends with griddepcontrol.wait, i.e. it completes only after the evaluation before it
// has completed -- what follows on the stream sees both.
struct PeerWaitArgs {
    const double* records;         // [world][2][batch] x 2 of the slot (local buffer)
    double* results;               // [world][2][batch] of the slot
    double* const* peer_base;      // for the acknowledgement
    long long ack_off;             // offset (doubles) of this rank's ack entry of the slot
    unsigned long long key, epoch;
    int n, world;
    int max_sleep_ns;              // polling back-off ceiling (the overlapped consumer shares the GPU with an evaluation)
    int* err;
};

__global__ void __launch_bounds__(1024) ctd_peer_wait_kernel(const PeerWaitArgs a) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    constexpr int U = 8;
    const int T = (int)blockDim.x;
    unsigned int sleep_ns = 100;
    for (int base = 0; base < a.n; base += T * U) {
        unsigned int pend = 0;
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (base + u * T + (int)threadIdx.x < a.n) pend |= 1u << u;
        while (pend) {
            unsigned long long b[U], chk[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {   // all loads of a pass in flight together
                const int q = base + u * T + (int)threadIdx.x;
                const double* r = a.records + 2 * (long long)(q < a.n ? q : 0);
                asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(b[u]), "=l"(chk[u]) : "l"(r) : "memory");
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if ((pend >> u & 1u) && (b[u] ^ chk[u]) == a.key) {
                    a.results[base + u * T + threadIdx.x] = __longlong_as_double((long long)b[u]);
                    pend &= ~(1u << u);
                }
            }
            if (pend) {
                __nanosleep(sleep_ns);
                if ((int)sleep_ns < a.max_sleep_ns) sleep_ns *= 2;
                long long t1;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
                if (t1 - t0 > 5000000000ll) { *a.err = 1; break; }   // 5 s: a rank never pushed (reported by ctd_sync)
            }
        }
    }
    __syncthreads();
    if ((int)threadIdx.x < a.world) {
        __threadfence_system();
        unsigned long long* ak = reinterpret_cast<unsigned long long*>(a.peer_base[threadIdx.x] + a.ack_off);
        asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(ak), "l"(a.epoch) : "memory");
    }
    asm volatile("griddepcontrol.wait;" ::: "memory");
}

const double* Model::peer_wait(int slot, bool overlapped) {
    if (!peer_local_) throw CtdError{-2, "ctd_peer_wait: call ctd_peer_init first"};
    if (slot < 0 || slot >= kCommSlots) throw CtdError{-2, "ctd_peer_wait: bad slot"};
    CTD_CUDA(cudaSetDevice(device_));
    const size_t per_slot = (size_t)comm_world_ * 2 * batch_;
    PeerWaitArgs w;
    w.records = peer_local_ + (size_t)slot * per_slot * 2;
    w.results = peer_local_ + peer_records() + (size_t)slot * per_slot;
    w.peer_base = d_peer_base_;
    w.ack_off = (long long)(peer_records() + peer_results()) + (long long)slot * comm_world_ + comm_rank_;
    w.epoch = peer_epoch_[slot];
    w.key = peer_key_of(w.epoch, slot);
    w.n = (int)per_slot;
    w.world = comm_world_;
    w.err = d_peer_err_;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(1, 1, 1);
    // in the evaluation stream the consumer is on the critical path (many threads, tight polling); beside an evaluation
    // it is a small CTA that backs off
    cfg.blockDim = dim3(overlapped ? 128 : (per_slot >= 4096 ? 1024 : 256), 1, 1);
    w.max_sleep_ns = overlapped ? 3200 : 100;
    cudaLaunchAttribute attr[1];
    if (overlapped) {
        // on the handle's communication stream, with NO dependency on the evaluation stream: the records are keyed by
        // the launch that pushes them, so the consumer may start polling at any time, and the evaluation stream holds
        // nothing but evaluation kernels (their programmatic dependent launches stay back to back); ctd_comm_wait(slot)
        // joins the consumer
        cfg.stream = comm_stream_;
        CTD_CUDA(cudaLaunchKernelEx(&cfg, ctd_peer_wait_kernel, w));
        CTD_CUDA(cudaEventRecord(ev_gathered_[slot], comm_stream_));
        slot_used_[slot] = true;
    } else {
        cfg.stream = stream_;
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        CTD_CUDA(cudaLaunchKernelEx(&cfg, ctd_peer_wait_kernel, w));
    }
    ++launches_;
    return w.results;
}

void Model::comm_destroy() {
    peer_destroy();
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
