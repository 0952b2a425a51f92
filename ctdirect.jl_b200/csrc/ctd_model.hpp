// Host-side model of one discretised OCP (the DOCP struct of the reference, src/DOCP_data.jl:265-366,
// plus what ADNLPModels keeps around it: sparsity structures and value buffers).
// Non-templated: everything problem/scheme specific comes from ElemInfo and the Launcher that the
// generated translation units register (ctd_model_impl.cuh).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "ctd_elem_common.cuh"
#include "ctd_kernels.cuh"

namespace ctd {

struct LaunchCfg {
    int threads = 0, own = 0, nred = 0, wpc = 1;
    size_t warp_smem = 0;   // bytes of one warp's staging region
    size_t tile_smem = 0;
    int map_step = 0, map_fp = 0, map_bc = 0;  // map sizes
    // "parts" variant for wide elements on small grids (0 threads: not compiled for this element)
    int p_threads = 0, p_own = 0;
    size_t p_smem = 0;
};

struct Launcher {
    // returns cudaError_t as int
    int (*launch)(const EvalArgs& a, int flags, int batch, size_t smem, cudaStream_t st, int parts, int threads) = nullptr;
    LaunchCfg (*cfg)(int flags) = nullptr;
    int (*prepare)(int flags, size_t smem, int parts) = nullptr;
    int (*unpack)(const UnpackArgs& a, int batch, cudaStream_t st) = nullptr;
    int (*pack)(const PackArgs& a, int batch, cudaStream_t st) = nullptr;  // opt in to > 48 KB dynamic shared memory
};

struct Registration {
    ElemInfo info;
    Launcher launcher;
};

void register_element(const Registration& r);
const std::vector<Registration>& registry();

// CSC sparsity structure, 1-based row indices (as SparseArrays.sparse returns it)
struct Csc {
    int64_t m = 0, n = 0;
    std::vector<int64_t> colptr;  // n+1, 0-based offsets
    std::vector<int64_t> rows;    // 1-based
    int64_t nnz() const { return (int64_t)rows.size(); }
    int64_t find(int64_t r, int64_t c) const;  // 1-based (r,c) -> 0-based position or -1
};

// rectangles -> CSC without a global sort (per-column interval lists)
class BlockPattern {
  public:
    void add(int64_t i0, int64_t i1, int64_t j0, int64_t j1, bool sym = false);
    Csc build(int64_t m, int64_t n, bool lower) const;

  private:
    struct Rect { int64_t i0, i1, j0, j1; };
    std::vector<Rect> rects_;
};

constexpr int kCommSlots = 4;      // result slots of the pipelined gather
constexpr int kNcclIdBytes = 128;  // sizeof(ncclUniqueId)

class Model {
  public:
    Model(const Registration& reg, int grid_size, const double* time_grid, int n_time_grid, const double* params,
          int n_params, int batch, int device);
    ~Model();

    const ElemInfo& info() const { return reg_.info; }
    int N() const { return N_; }
    int batch() const { return batch_; }
    int64_t nvar() const { return nvar_; }
    int64_t ncon() const { return ncon_; }
    int64_t nnzj() const { return jac_.nnz(); }
    int64_t nnzh() const { return hess_.nnz(); }
    const Csc& jac() const { return jac_; }
    const Csc& hess() const { return hess_; }
    const std::vector<double>& sgrid() const { return sgrid_; }

    void variable_bounds(double* lvar, double* uvar) const;    // DOCP_variables.jl:21-63
    void constraint_bounds(double* lcon, double* ucon) const;  // DOCP_functions.jl:163-191
    void initial_guess(double* x0) const;                      // DOCP_variables.jl:122-145
    void structure(const Csc& p, int64_t* rows, int64_t* cols) const;

    // device evaluation; pointers are DEVICE pointers, arrays are [batch][dim]
    // `viol` (optional, needs F_CONS): per-instance max(lcon - c, c - ucon, 0), reduced in the same launch
    void eval_dev(int flags, const double* x, const double* y, double sigma, double* obj, double* grad, double* c,
                  double* jvals, double* hvals, double* viol = nullptr);
    // host-pointer evaluation (copies through the handle's device buffers)
    void eval_host(int flags, const double* x, const double* y, double sigma, double* obj, double* grad, double* c,
                   double* jvals, double* hvals, bool wait = true);
    void norms_dev(const double* x, double* obj, double* viol);
    void violation_dev(const double* c, double* viol);
    void norms_host(const double* x, double* obj, double* viol);
    // solution unpacking (device pointers; any output may be NULL)
    void unpack_dev(const double* x, const double* y, double* time, double* state, double* control, double* variable,
                    double* costate, double* path_dual, double* boundary_dual);
    void unpack_host(const double* x, const double* y, double* time, double* state, double* control, double* variable,
                     double* costate, double* path_dual, double* boundary_dual);
    // initial guess / warm start from sampled data (device or host pointers; inputs may be NULL)
    void pack_dev(const double* state, const double* control, const double* variable, const double* stage_control, double* x);
    void pack_host(const double* state, const double* control, const double* variable, const double* stage_control, double* x);
    void set_params(const double* params, int n_params);
    void sync();
    cudaStream_t stream();
    // run this handle's work on a caller-owned stream (nullptr: back to the handle's own stream)
    void set_stream(cudaStream_t s);
    int64_t launches() const { return launches_; }
    // KKT assembly (ctd_kkt.cu)
    int64_t kkt_nnz();
    void kkt_structure(int64_t* colptr, int64_t* rows);
    void kkt_assemble_dev(const double* hvals, const double* jvals, const double* dx, const double* dc, double* kvals);
    // multi-GPU norm gather (ctd_comm.cu)
    void comm_init(const char* nccl_id, int rank, int world);
    void comm_destroy();
    void gather_norms_dev(const double* local, double* gathered, int slot);
    void comm_wait(int slot);
    int comm_world() const { return comm_world_; }
    // fused norm gather over peer memory (ctd_comm.cu)
    void peer_init();
    void eval_gather_dev(int flags, const double* x, const double* y, double sigma, double* grad, double* c, double* jvals,
                         double* hvals, int slot);
    const double* peer_wait(int slot, bool overlapped);   // enqueues the consumer; returns the [world][2][batch] rows of the slot

  private:
    void build_patterns();
    void build_detected_patterns();
    void build_edge_tables();
    void ensure_device();
    void ensure_staging();
    void ensure_con_bounds();
    EvalArgs base_args() const;

    Registration reg_;
    int N_ = 0, batch_ = 1, device_ = 0;
    std::vector<double> sgrid_, params_;
    int64_t nvar_ = 0, ncon_ = 0;
    Csc jac_, hess_;
    // edge tables (host copies)
    std::vector<int> edge_map_;
    std::vector<int64_t> edge_pos_;
    int ne_j_ = 0, ne_h_ = 0, ne_g_ = 0, ntask_ = 0;
    EdgeTask task_[MAX_TASKS];
    int fp_map_off_ = 0, bc_map_off_ = 0, ip_map_off_ = 0;
    int64_t jv_base_[MAX_NV], hvv_pos_[MAX_NVV];
    // device state (created on first evaluation)
    bool dev_ready_ = false, staging_ready_ = false;
    cudaStream_t stream_ = nullptr, own_stream_ = nullptr;
    int* d_edge_map_ = nullptr;
    long long* d_edge_pos_ = nullptr;
    double *d_sgrid_ = nullptr, *d_params_ = nullptr, *d_partials_ = nullptr;
    unsigned int* d_counters_ = nullptr;
    double *d_x_ = nullptr, *d_y_ = nullptr, *d_c_ = nullptr, *d_j_ = nullptr, *d_h_ = nullptr, *d_g_ = nullptr,
           *d_obj_ = nullptr, *d_lcon_ = nullptr, *d_ucon_ = nullptr, *d_viol_ = nullptr;
    int64_t launches_ = 0;
    // KKT structure and position maps (built on first use)
    void build_kkt();
    void kkt_free();
    bool kkt_built_ = false;
    std::vector<int64_t> kkt_colptr_, kkt_rows_, kkt_pos_h_, kkt_pos_j_, kkt_hdiag_;
    int *d_kkt_pos_h_ = nullptr, *d_kkt_pos_j_ = nullptr, *d_kkt_hdiag_ = nullptr, *d_kkt_colptr_ = nullptr;
    // communicator state (only after comm_init)
    void* comm_ = nullptr;
    int comm_rank_ = 0, comm_world_ = 0;
    cudaStream_t comm_stream_ = nullptr;
    cudaEvent_t ev_issue_ = nullptr, ev_gathered_[kCommSlots] = {};
    bool slot_used_[kCommSlots] = {};
    // peer-memory gather
    struct PeerPush { int slot; unsigned long long epoch; };
    void peer_destroy();
    size_t peer_records() const;
    size_t peer_results() const;
    void peer_args(EvalArgs& a) const;
    double* peer_local_ = nullptr;            // this rank's gather buffer (values, then arrival flags)
    std::vector<double*> peer_base_;          // every rank's buffer as mapped here
    double** d_peer_base_ = nullptr;
    unsigned long long* d_peer_gate_ = nullptr;
    int* d_peer_err_ = nullptr;
    unsigned long long peer_epoch_[kCommSlots] = {};
    const PeerPush* peer_push_ = nullptr;     // set while eval_gather_dev builds its launch
};

void nccl_unique_id(char* id);   // kNcclIdBytes bytes
int nccl_version();

struct CtdError {
    int code;
    std::string msg;
};

}  // namespace ctd
