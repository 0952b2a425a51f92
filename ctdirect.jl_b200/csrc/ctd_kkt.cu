// KKT assembly: the lower triangle of the augmented system an interior-point solver
// factorises every iteration,
//        K = [ H + diag(dx)       .      ]        H: Lagrangian Hessian (tril pattern of the model)
//            [ J              -diag(dc)  ]        J: constraint Jacobian
// in CSC order (column-major, rows ascending, 1-based Int64 like the other structures).  In the reference
// this matrix is assembled by the solver from the hess_coord! / jac_coord! arrays (MadNLP's sparse KKT
// system, then cuDSS: test/test_gpu.jl:18); here the structure and the position maps are built once on the
// host and one kernel moves the values the evaluation kernels left in HBM to their KKT positions.
//
// Column j <= nvar: diagonal entry (always present, also when the Hessian pattern has no (j,j) entry),
// the strictly lower Hessian entries of column j, then the Jacobian entries of column j at rows nvar + r.
// Column nvar + r: its diagonal entry.  Because both patterns are already CSC, the KKT value array is a
// column-wise interleaving of hvals and jvals: no sort is involved.
#include "ctd_model.hpp"

namespace ctd {

#define CTD_CUDA(call)                                                                         \
    do {                                                                                       \
        cudaError_t _e = (call);                                                               \
        if (_e != cudaSuccess)                                                                 \
            throw CtdError{(int)_e, std::string("CUDA error: ") + cudaGetErrorString(_e) + " at " #call}; \
    } while (0)

void Model::build_kkt() {
    if (kkt_built_) return;
    const int64_t n = nvar_, m = ncon_;
    kkt_colptr_.assign(n + m + 1, 0);
    kkt_rows_.clear();
    kkt_rows_.reserve((size_t)(hess_.nnz() + jac_.nnz() + n + m));
    kkt_pos_h_.assign((size_t)hess_.nnz(), -1);
    kkt_pos_j_.assign((size_t)jac_.nnz(), -1);
    kkt_hdiag_.assign((size_t)n, -1);
    for (int64_t j = 0; j < n; ++j) {
        kkt_colptr_[j] = (int64_t)kkt_rows_.size();
        kkt_rows_.push_back(j + 1);   // diagonal
        for (int64_t k = hess_.colptr[j]; k < hess_.colptr[j + 1]; ++k) {
            const int64_t r = hess_.rows[k];   // 1-based, >= j + 1 (lower triangle)
            if (r == j + 1) {
                kkt_hdiag_[j] = k;             // folded into the diagonal entry by the kernel
            } else {
                kkt_pos_h_[k] = (int64_t)kkt_rows_.size();
                kkt_rows_.push_back(r);
            }
        }
        for (int64_t k = jac_.colptr[j]; k < jac_.colptr[j + 1]; ++k) {
            kkt_pos_j_[k] = (int64_t)kkt_rows_.size();
            kkt_rows_.push_back(n + jac_.rows[k]);
        }
    }
    for (int64_t r = 0; r < m; ++r) {
        kkt_colptr_[n + r] = (int64_t)kkt_rows_.size();
        kkt_rows_.push_back(n + r + 1);
    }
    kkt_colptr_[n + m] = (int64_t)kkt_rows_.size();
    kkt_built_ = true;
}

int64_t Model::kkt_nnz() {
    build_kkt();
    return (int64_t)kkt_rows_.size();
}

void Model::kkt_structure(int64_t* colptr, int64_t* rows) {
    build_kkt();
    if (colptr)
        for (size_t j = 0; j < kkt_colptr_.size(); ++j) colptr[j] = kkt_colptr_[j] + 1;   // 1-based like Julia's colptr
    if (rows) std::copy(kkt_rows_.begin(), kkt_rows_.end(), rows);
}

struct KktArgs {
    const double *hvals, *jvals, *dx, *dc;
    double* kvals;
    const int *pos_h, *pos_j, *hdiag, *diag;   // 32-bit positions inside one instance (nnzk < 2^31 is checked)
    long long nnzh, nnzj, nvar, ncon, nnzk;
};

// One grid-stride pass per source array.  Diagonal Hessian entries are skipped by the first section
// (pos_h < 0) and written, together with the regularisation, by the third: no two threads write one entry.
__global__ void __launch_bounds__(256) ctd_kkt_kernel(const KktArgs a) {
    const long long inst = blockIdx.y;
    const double* __restrict__ hv = a.hvals + inst * a.nnzh;
    const double* __restrict__ jv = a.jvals + inst * a.nnzj;
    double* __restrict__ kv = a.kvals + inst * a.nnzk;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long t0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (long long k = t0; k < a.nnzh; k += stride) {
        const int p = __ldg(a.pos_h + k);
        if (p >= 0) kv[p] = hv[k];
    }
    for (long long k = t0; k < a.nnzj; k += stride) kv[__ldg(a.pos_j + k)] = jv[k];
    for (long long j = t0; j < a.nvar + a.ncon; j += stride) {
        double v;
        if (j < a.nvar) {
            const int d = __ldg(a.hdiag + j);
            v = (d >= 0 ? hv[d] : 0.0) + (a.dx ? a.dx[inst * a.nvar + j] : 0.0);
        } else {
            v = a.dc ? -a.dc[inst * a.ncon + (j - a.nvar)] : 0.0;
        }
        kv[__ldg(a.diag + j)] = v;
    }
}

void Model::kkt_assemble_dev(const double* hvals, const double* jvals, const double* dx, const double* dc, double* kvals) {
    if (!hvals || !jvals || !kvals) throw CtdError{-2, "ctd_kkt_assemble_dev: NULL buffer"};
    build_kkt();
    ensure_device();
    CTD_CUDA(cudaSetDevice(device_));
    if (!d_kkt_pos_h_) {
        if (kkt_rows_.size() >= (size_t)1 << 31) throw CtdError{-2, "ctd_kkt_assemble_dev: more than 2^31 KKT entries per instance"};
        auto up = [&](int** d, const std::vector<int64_t>& h, size_t n) {
            std::vector<int> h32(h.begin(), h.begin() + n);
            CTD_CUDA(cudaMalloc(d, sizeof(int) * std::max<size_t>(n, 1)));
            CTD_CUDA(cudaMemcpyAsync(*d, h32.data(), sizeof(int) * n, cudaMemcpyHostToDevice, stream_));
            CTD_CUDA(cudaStreamSynchronize(stream_));   // h32 is a temporary
        };
        up(&d_kkt_pos_h_, kkt_pos_h_, kkt_pos_h_.size());
        up(&d_kkt_pos_j_, kkt_pos_j_, kkt_pos_j_.size());
        up(&d_kkt_hdiag_, kkt_hdiag_, kkt_hdiag_.size());
        up(&d_kkt_colptr_, kkt_colptr_, kkt_colptr_.size() - 1);   // every column starts on its diagonal entry
    }
    KktArgs a;
    a.hvals = hvals; a.jvals = jvals; a.dx = dx; a.dc = dc; a.kvals = kvals;
    a.pos_h = d_kkt_pos_h_; a.pos_j = d_kkt_pos_j_; a.hdiag = d_kkt_hdiag_; a.diag = d_kkt_colptr_;
    a.nnzh = hess_.nnz(); a.nnzj = jac_.nnz(); a.nvar = nvar_; a.ncon = ncon_; a.nnzk = (long long)kkt_rows_.size();
    const long long work = std::max<long long>(a.nnzh, std::max<long long>(a.nnzj, a.nvar + a.ncon));
    // a multiple of the SM count, 8 CTAs of 256 threads per SM at most, never more CTAs than work
    long long ctas = std::min<long long>((work + 255) / 256, (long long)148 * 8);
    if (batch_ > 1) ctas = std::min<long long>(ctas, std::max<long long>(1, (long long)148 * 8 / batch_));
    dim3 grid((unsigned)std::max<long long>(ctas, 1), (unsigned)batch_, 1);
    ctd_kkt_kernel<<<grid, 256, 0, stream_>>>(a);
    CTD_CUDA(cudaGetLastError());
    ++launches_;
}

void Model::kkt_free() {
    cudaFree(d_kkt_pos_h_); cudaFree(d_kkt_pos_j_); cudaFree(d_kkt_hdiag_); cudaFree(d_kkt_colptr_);
    d_kkt_pos_h_ = d_kkt_pos_j_ = d_kkt_hdiag_ = d_kkt_colptr_ = nullptr;
}

}  // namespace ctd
