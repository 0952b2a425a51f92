// Host-side model: layout, bounds, initial guess, sparsity structures, edge tables and kernel launches.
// Reference functions followed (paths relative to the CTDirect.jl tree):
//   DOCP constructor / scheme dims   src/DOCP_data.jl:176-214,293-365; trapeze.jl:14-42; midpoint.jl:17-39;
//                                    irk.jl:138-160; irk_stagewise.jl:136-163
//   bounds / initial guess           src/DOCP_functions.jl:163-191; src/DOCP_variables.jl:21-145;
//                                    irk_stagewise.jl:250-335
//   sparsity patterns                trapeze.jl:149-303; midpoint.jl:163-300; irk.jl:315-496;
//                                    irk_stagewise.jl:468-638; common.jl:285-312
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>

#include "ctd_model.hpp"

#include <mutex>
#include <tuple>

namespace ctd {

#define CTD_CUDA(call)                                                                         \
    do {                                                                                       \
        cudaError_t _e = (call);                                                               \
        if (_e != cudaSuccess)                                                                 \
            throw CtdError{(int)_e, std::string("CUDA error: ") + cudaGetErrorString(_e) + " at " #call}; \
    } while (0)

static std::vector<Registration>& registry_mut() {
    static std::vector<Registration> r;
    return r;
}
void register_element(const Registration& r) { registry_mut().push_back(r); }
const std::vector<Registration>& registry() { return registry_mut(); }

// --------------------------------------------------------------------------------------------
int64_t Csc::find(int64_t r, int64_t c) const {
    if (c < 1 || c > n) return -1;
    const int64_t* b = rows.data() + colptr[c - 1];
    const int64_t* e = rows.data() + colptr[c];
    const int64_t* it = std::lower_bound(b, e, r);
    if (it == e || *it != r) return -1;
    return it - rows.data();
}

void BlockPattern::add(int64_t i0, int64_t i1, int64_t j0, int64_t j1, bool sym) {
    if (i1 < i0 || j1 < j0) return;  // empty Julia range: add_nonzero_block! loops do nothing
    rects_.push_back({i0, i1, j0, j1});
    if (sym) rects_.push_back({j0, j1, i0, i1});
}

// Equivalent of SparseArrays.sparse(Is, Js, trues, m, n) (CSC, rows ascending inside a column,
// duplicates merged) followed, for the Hessian, by tril: every rectangle contributes one row interval
// to each of its columns; per column the intervals are sorted and merged.
Csc BlockPattern::build(int64_t m, int64_t n, bool lower) const {
    std::vector<int64_t> cnt(n + 1, 0);
    for (const Rect& r : rects_)
        for (int64_t j = r.j0; j <= r.j1; ++j) cnt[j]++;
    std::vector<int64_t> start(n + 2, 0);
    for (int64_t j = 1; j <= n; ++j) start[j + 1] = start[j] + cnt[j];
    std::vector<std::pair<int64_t, int64_t>> iv(start[n + 1]);
    std::vector<int64_t> fill(start.begin(), start.end());
    for (const Rect& r : rects_)
        for (int64_t j = r.j0; j <= r.j1; ++j) iv[fill[j]++] = {r.i0, r.i1};
    Csc p;
    p.m = m;
    p.n = n;
    p.colptr.assign(n + 1, 0);
    // first pass: count, second pass: emit
    for (int pass = 0; pass < 2; ++pass) {
        int64_t pos = 0;
        for (int64_t j = 1; j <= n; ++j) {
            auto b = iv.begin() + start[j], e = iv.begin() + start[j + 1];
            if (pass == 0) std::sort(b, e);
            int64_t covered = lower ? j - 1 : 0;  // rows <= covered are already emitted / excluded
            for (auto it = b; it != e; ++it) {
                int64_t lo = std::max(it->first, covered + 1), hi = it->second;
                if (hi < lo) continue;
                if (pass == 1)
                    for (int64_t r = lo; r <= hi; ++r) p.rows[pos + (r - lo)] = r;
                pos += hi - lo + 1;
                covered = hi;
            }
            if (pass == 0) p.colptr[j] = pos;
        }
        if (pass == 0) p.rows.resize(pos);
    }
    return p;
}

// --------------------------------------------------------------------------------------------
Model::Model(const Registration& reg, int grid_size, const double* time_grid, int n_time_grid,
             const double* params, int n_params, int batch, int device)
    : reg_(reg), batch_(batch), device_(device) {
    const ElemInfo& f = reg_.info;
    if (batch < 1) throw CtdError{-2, "batch must be >= 1"};
    if (batch > 65535) throw CtdError{-2, "batch must be <= 65535 (instances are the y dimension of the grid)"};
    if (f.nv > MAX_NV) throw CtdError{-2, "too many optimisation variables v for this build"};
    // DOCPtime: DOCP_data.jl:176-214
    if (!time_grid) {
        if (grid_size < 1) throw CtdError{-2, "grid_size must be >= 1"};
        N_ = grid_size;
        sgrid_.resize(N_ + 1);
        for (int j = 0; j <= N_; ++j) sgrid_[j] = (double)j / (double)N_;  // collect(LinRange(0, 1, N+1))
    } else {
        if (n_time_grid < 2) throw CtdError{-2, "time grid needs at least 2 points"};
        for (int j = 0; j + 1 < n_time_grid; ++j)
            if (!(time_grid[j] < time_grid[j + 1]))
                throw CtdError{-3, "given time grid is not strictly increasing. Aborting..."};  // DOCP_data.jl:187
        N_ = n_time_grid - 1;
        sgrid_.assign(time_grid, time_grid + n_time_grid);
        if (time_grid[0] != 0.0 || time_grid[N_] != 1.0) {  // normalisation DOCP_data.jl:191-196
            const double a = time_grid[0], b = time_grid[N_];
            for (int j = 0; j <= N_; ++j) sgrid_[j] = (time_grid[j] - a) / (b - a);
        }
    }
    nvar_ = (int64_t)N_ * f.blk + f.nx + f.nv + (f.final_control ? f.nu : 0);
    ncon_ = (int64_t)N_ * f.cb + f.np + f.nb;
    // parameters: one row per instance
    const int npar = std::max(f.npar, 1);
    params_.assign((size_t)batch_ * npar, 0.0);
    for (int b = 0; b < batch_; ++b)
        for (int k = 0; k < f.npar; ++k) params_[(size_t)b * npar + k] = f.meta.default_params[k];
    if (params && n_params > 0) {
        if (n_params == f.npar) {
            for (int b = 0; b < batch_; ++b)
                for (int k = 0; k < f.npar; ++k) params_[(size_t)b * npar + k] = params[k];
        } else if (n_params == f.npar * batch_) {
            for (int b = 0; b < batch_; ++b)
                for (int k = 0; k < f.npar; ++k) params_[(size_t)b * npar + k] = params[(size_t)b * f.npar + k];
        } else {
            throw CtdError{-2, "params must hold npar or batch*npar values"};
        }
    }
    build_patterns();
    build_edge_tables();
}

Model::~Model() {
    if (dev_ready_) {
        cudaSetDevice(device_);
        comm_destroy();
        kkt_free();
        if (stream_) cudaStreamSynchronize(stream_);   // nothing of this handle may still be in flight
        cudaFree(d_edge_map_); cudaFree(d_edge_pos_); cudaFree(d_sgrid_); cudaFree(d_params_);
        cudaFree(d_partials_); cudaFree(d_counters_);
        cudaFree(d_x_); cudaFree(d_y_); cudaFree(d_c_); cudaFree(d_j_); cudaFree(d_h_); cudaFree(d_g_); cudaFree(d_obj_);
        cudaFree(d_lcon_); cudaFree(d_ucon_); cudaFree(d_viol_);
        if (own_stream_) cudaStreamDestroy(own_stream_);
    }
}

// --------------------------------------------------------------------------------------------
// ":optimized" (detected) patterns, pattern mode 2: assembled from the structure lists of the three
// element kinds that the generator obtained by tracing the OCP expressions (gen/trace.py); the entries are
// gathered as (column, row) keys, sorted and made unique -- the CSC order of SparseArrays.sparse.
static Csc csc_from_keys(std::vector<std::pair<int64_t, int64_t>>& key, int64_t m, int64_t n) {
    std::sort(key.begin(), key.end());
    key.erase(std::unique(key.begin(), key.end()), key.end());
    Csc p;
    p.m = m;
    p.n = n;
    p.colptr.assign(n + 1, 0);
    p.rows.resize(key.size());
    for (size_t k = 0; k < key.size(); ++k) {
        p.rows[k] = key[k].second;
        p.colptr[key[k].first]++;
    }
    for (int64_t j = 0; j < n; ++j) p.colptr[j + 1] += p.colptr[j];
    return p;
}

void Model::build_detected_patterns() {
    const ElemInfo& f = reg_.info;
    const int64_t B = f.blk, CB = f.cb, N = N_, nv = f.nv, nloc = f.nown + f.nnext;
    auto scol = [&](int64_t i, int a) -> int64_t { return a < nloc ? (i - 1) * B + a + 1 : nvar_ - nv + (a - nloc) + 1; };
    auto fcol = [&](int a) -> int64_t {   // final-path element: (u_f block, x_f, v), or (x_f, u_f, v) for trapeze
        const int nl = f.fp_nloc - f.nv;
        if (a >= nl) return nvar_ - nv + (a - nl) + 1;
        if (f.final_control) return N * B + a + 1;
        if (a < f.ctrl_blk) return (N - 1) * B + f.nx + a + 1;
        return N * B + (a - f.ctrl_blk) + 1;
    };
    auto bcol = [&](int a) -> int64_t {   // boundary element: (x_1, x_{N+1}, v)
        if (a < f.nx) return a + 1;
        if (a < 2 * f.nx) return N * B + (a - f.nx) + 1;
        return nvar_ - nv + (a - 2 * f.nx) + 1;
    };
    std::vector<std::pair<int64_t, int64_t>> J, H;
    J.reserve((size_t)N * f.s_step_j.n + 64);
    H.reserve((size_t)N * f.s_step_h.n + 64);
    auto addH = [&](int64_t ga, int64_t gb) { H.emplace_back(std::min(ga, gb), std::max(ga, gb)); };   // (col, row), row >= col
    for (int64_t i = 1; i <= N; ++i) {
        for (int n = 0; n < f.s_step_j.n; ++n) {
            const int r = f.s_step_j(n, 0);
            if (f.path_shift && r >= f.eqs) {   // path rows of node i+1 (rows of the next step's block); node N+1: final-path element
                if (i < N) J.emplace_back(scol(i, f.s_step_j(n, 1)), i * CB + r + 1);
                continue;
            }
            J.emplace_back(scol(i, f.s_step_j(n, 1)), (i - 1) * CB + r + 1);
        }
        for (int n = 0; n < f.s_step_h.n; ++n) addH(scol(i, f.s_step_h(n, 0)), scol(i, f.s_step_h(n, 1)));
    }
    if (f.path_shift) {   // path rows of node 1: the final-path structure on (U_1 block, x_1, v)
        const int nl = f.fp_nloc - f.nv;
        auto icol = [&](int a) -> int64_t {
            if (a >= nl) return nvar_ - nv + (a - nl) + 1;
            if (a < f.ctrl_blk) return f.nx + a + 1;
            return (a - f.ctrl_blk) + 1;
        };
        for (int n = 0; n < f.s_fp_j.n; ++n) J.emplace_back(icol(f.s_fp_j(n, 1)), f.eqs + f.s_fp_j(n, 0) + 1);
        for (int n = 0; n < f.s_fp_h.n; ++n) addH(icol(f.s_fp_h(n, 0)), icol(f.s_fp_h(n, 1)));
    }
    for (int n = 0; n < f.s_fp_j.n; ++n) J.emplace_back(fcol(f.s_fp_j(n, 1)), N * CB + f.s_fp_j(n, 0) + 1);
    for (int n = 0; n < f.s_fp_h.n; ++n) addH(fcol(f.s_fp_h(n, 0)), fcol(f.s_fp_h(n, 1)));
    for (int n = 0; n < f.s_bc_j.n; ++n) J.emplace_back(bcol(f.s_bc_j(n, 1)), ncon_ - f.nb + f.s_bc_j(n, 0) + 1);
    for (int n = 0; n < f.s_bc_h.n; ++n) addH(bcol(f.s_bc_h(n, 0)), bcol(f.s_bc_h(n, 1)));
    jac_ = csc_from_keys(J, ncon_, nvar_);
    hess_ = csc_from_keys(H, nvar_, nvar_);
}

void Model::build_patterns() {
    const ElemInfo& f = reg_.info;
    if (f.pattern_mode == 2) {
        build_detected_patterns();
        return;
    }
    const int64_t nx = f.nx, nu = f.nu, nv = f.nv, np = f.np, nb = f.nb, B = f.blk, CB = f.cb, N = N_;
    const int64_t s = f.stage, cblk = f.ctrl_blk;
    const int64_t v0 = nvar_ - nv + 1, v1 = nvar_;
    const bool corrected = f.pattern_mode == 1;
    const bool trapeze = f.scheme == "trapeze", midpoint = f.scheme == "midpoint";
    const bool euler_e = f.scheme == "euler", euler_i = f.scheme == "euler_implicit";
    BlockPattern J, H;
    H.add(v0, v1, v0, v1);
    for (int64_t i = 1; i <= N; ++i) {
        const int64_t co = (i - 1) * CB, vo = (i - 1) * B;
        const int64_t dyn0 = co + 1, dyn1 = co + nx;
        if (midpoint) {
            J.add(dyn0, dyn1, vo + 1, vo + B + nx);             // state eq: x_i, u_i, x_{i+1}
            J.add(co + nx + 1, co + CB, vo + 1, vo + nx + nu);  // path: x_i, u_i
            J.add(co + 1, co + CB, v0, v1);                     // everything depends on v
            H.add(vo + 1, vo + B + nx, vo + 1, vo + B + nx);
            H.add(vo + 1, vo + B + nx, v0, v1, true);
        } else if (euler_e || euler_i) {   // euler.jl:166-355
            const int64_t p0 = co + nx + 1, p1 = co + CB;
            const int64_t u0 = vo + nx + 1, u1 = vo + nx + nu;
            J.add(dyn0, dyn1, vo + 1, vo + nx);
            J.add(dyn0, dyn1, u0, vo + B + nx);
            J.add(dyn0, dyn1, v0, v1);
            J.add(p0, p1, vo + 1, vo + nx);
            if (euler_i && corrected && i >= 2)
                J.add(p0, p1, u0 - B, u1 - B);   // node i reads the control of block i-1 (euler.jl:66-69)
            else
                J.add(p0, p1, u0, u1);
            J.add(p0, p1, v0, v1);
            H.add(u0, u1, u0, u1);
            H.add(u0, u1, v0, v1, true);
            if (euler_i) {
                H.add(vo + B + 1, vo + B + nx, vo + B + 1, vo + B + nx);
                H.add(vo + B + 1, vo + B + nx, u0, u1, true);
                H.add(vo + B + 1, vo + B + nx, v0, v1, true);
            }
            H.add(vo + 1, vo + nx, vo + 1, vo + nx);
            H.add(vo + 1, vo + nx, u0, u1, true);
            H.add(vo + 1, vo + nx, v0, v1, true);
        } else if (trapeze) {
            const int64_t p0 = co + nx + 1, p1 = co + CB;
            J.add(dyn0, dyn1, vo + 1, vo + nx);
            J.add(dyn0, dyn1, vo + nx + 1, vo + nx + nu + nx);
            J.add(dyn0, dyn1, vo + 2 * nx + nu + 1, vo + 2 * nx + 2 * nu);
            J.add(p0, p1, vo + 1, vo + nx);
            J.add(p0, p1, vo + nx + 1, vo + nx + nu);
            J.add(p0, p1, v0, v1);                         // trapeze.jl:203: path rows only
            if (corrected) J.add(dyn0, dyn1, v0, v1);      // the h_i(tf) dependency the reference omits
            H.add(vo + 1, vo + 2 * B, vo + 1, vo + 2 * B);
            H.add(vo + 1, vo + 2 * B, v0, v1, true);
        } else {  // Gauss-Legendre IRK, constant or stagewise controls
            const int64_t st0 = co + nx + 1, st1 = co + (s + 1) * nx;
            const int64_t p0 = co + (s + 1) * nx + 1, p1 = co + CB;
            const int64_t xi0 = vo + 1, xi1 = vo + nx, ui1 = vo + nx + cblk;
            const int64_t ki0 = vo + nx + cblk + 1, ki1 = vo + B;
            J.add(dyn0, dyn1, xi0, xi1);
            J.add(dyn0, dyn1, ki0, vo + B + nx);
            J.add(dyn0, dyn1, v0, v1);
            J.add(st0, st1, xi0, ki1);
            J.add(st0, st1, v0, v1);
            J.add(p0, p1, xi0, ui1);
            J.add(p0, p1, v0, v1);
            H.add(xi0, ki1, xi0, ki1);
            H.add(xi0, ki1, v0, v1, true);
        }
    }
    {
        const int64_t co = N * CB, vo = N * B;
        const int64_t xf0 = vo + 1, xf1 = vo + nx;
        const int64_t uf0 = trapeze ? vo + nx + 1 : vo - B + nx + 1;
        const int64_t uf1 = trapeze ? vo + nx + nu : vo - B + nx + cblk;
        J.add(co + 1, co + np, xf0, xf1);   // final path constraints
        J.add(co + 1, co + np, uf0, uf1);
        J.add(co + 1, co + np, v0, v1);
        const int64_t bo = co + np;         // boundary constraints
        J.add(bo + 1, bo + nb, 1, nx);
        J.add(bo + 1, bo + nb, xf0, xf1);
        J.add(bo + 1, bo + nb, v0, v1);
        if ((f.stagewise || euler_e || euler_i) && f.lagrange && !corrected)
            J.add(ncon_, ncon_, nx, nx);  // irk_stagewise.jl:550-552, euler.jl:257-259
        if (f.irk) {
            H.add(xf0, xf1, xf0, xf1);
            H.add(uf0, uf1, uf0, uf1);
            H.add(xf0, xf1, uf0, uf1, true);
            if (f.stagewise) {
                H.add(xf0, uf1, v0, v1, true);  // irk_stagewise.jl:624: xf0 > uf1, empty range
                if (corrected) H.add(xf0, xf1, v0, v1, true);
            } else {
                H.add(xf0, xf1, v0, v1, true);
            }
            H.add(uf0, uf1, v0, v1, true);
            H.add(1, nx, xf0, xf1, true);
        } else if (midpoint || euler_e || euler_i) {
            H.add(1, nx, xf0, xf1, true);  // midpoint.jl:286-291, euler.jl:345-350: unconditional
        } else if (f.mayer || nb > 0) {
            H.add(1, nx, xf0, xf1, true);  // trapeze.jl:288-295
        }
    }
    jac_ = J.build(ncon_, nvar_, false);
    hess_ = H.build(nvar_, nvar_, true);
}

void Model::structure(const Csc& p, int64_t* rows, int64_t* cols) const {
    for (int64_t j = 0; j < p.n; ++j)
        for (int64_t k = p.colptr[j]; k < p.colptr[j + 1]; ++k) {
            rows[k] = p.rows[k];
            cols[k] = j + 1;
        }
}

// --------------------------------------------------------------------------------------------
static void box_to_dense(int dim, const std::vector<Box>& box, std::vector<double>& lb, std::vector<double>& ub) {
    lb.assign(dim, -CTD_INF);
    ub.assign(dim, CTD_INF);
    for (const Box& b : box) {
        lb[b.idx - 1] = b.lb;
        ub[b.idx - 1] = b.ub;
    }
}

void Model::variable_bounds(double* lvar, double* uvar) const {
    const ElemInfo& f = reg_.info;
    std::fill(lvar, lvar + nvar_, -CTD_INF);
    std::fill(uvar, uvar + nvar_, CTD_INF);
    std::vector<double> xl, xu, ul, uu, vl, vu;
    box_to_dense(f.nx, f.meta.state_box, xl, xu);
    box_to_dense(f.nu, f.meta.control_box, ul, uu);
    box_to_dense(f.nv, f.meta.variable_box, vl, vu);
    // one control per stage (irk_stagewise.jl:250-300) or per control step (DOCP_variables.jl:44)
    const int ncopies = f.stagewise ? f.stage : f.control_steps;
    for (int64_t i = 0; i <= N_; ++i) {
        const int64_t o = i * f.blk;
        for (int k = 0; k < f.nx; ++k) { lvar[o + k] = xl[k]; uvar[o + k] = xu[k]; }
        if (i < N_ || f.final_control)
            for (int j = 0; j < ncopies; ++j)
                for (int k = 0; k < f.nu; ++k) {
                    lvar[o + f.nx + j * f.nu + k] = ul[k];
                    uvar[o + f.nx + j * f.nu + k] = uu[k];
                }
    }
    for (int k = 0; k < f.nv; ++k) { lvar[nvar_ - f.nv + k] = vl[k]; uvar[nvar_ - f.nv + k] = vu[k]; }
}

void Model::constraint_bounds(double* lcon, double* ucon) const {
    const ElemInfo& f = reg_.info;
    std::fill(lcon, lcon + ncon_, 0.0);
    std::fill(ucon, ucon + ncon_, 0.0);
    for (int64_t i = 0; i <= N_; ++i) {
        const int64_t o = (i < N_) ? i * f.cb + f.eqs : (int64_t)N_ * f.cb;
        for (int k = 0; k < f.np; ++k) { lcon[o + k] = f.meta.path_lb[k]; ucon[o + k] = f.meta.path_ub[k]; }
    }
    for (int k = 0; k < f.nb; ++k) { lcon[ncon_ - f.nb + k] = f.meta.bc_lb[k]; ucon[ncon_ - f.nb + k] = f.meta.bc_ub[k]; }
}

void Model::initial_guess(double* x0) const {
    const ElemInfo& f = reg_.info;
    std::fill(x0, x0 + nvar_, 0.1);  // DOCP_variables.jl:126
    const int ncopies = f.stagewise ? f.stage : f.control_steps;   // DOCP_variables.jl:138
    for (int64_t i = 0; i <= N_; ++i) {
        const int64_t o = i * f.blk;
        if (!f.meta.init_state.empty())
            for (int k = 0; k < f.nx; ++k) x0[o + k] = f.meta.init_state[k];
        if (!f.meta.init_control.empty() && (i < N_ || f.final_control))
            for (int j = 0; j < ncopies; ++j)
                for (int k = 0; k < f.nu; ++k) x0[o + f.nx + j * f.nu + k] = f.meta.init_control[k];
    }
    if (!f.meta.init_variable.empty())
        for (int k = 0; k < f.nv; ++k) x0[nvar_ - f.nv + k] = f.meta.init_variable[k];
}

// --------------------------------------------------------------------------------------------
// Edge tables: every value that lives in column blocks 1, N, N+1 (or in the boundary / final-path
// rows of the v columns) gets an entry in the edge buffer; the maps send each value produced by an
// edge element to its entry.
void Model::build_edge_tables() {
    const ElemInfo& f = reg_.info;
    const LaunchCfg lc = reg_.launcher.cfg(F_CONS | F_JAC | F_HESS | F_GRAD | F_OBJ);
    const int64_t B = f.blk, CB = f.cb, N = N_, nv = f.nv;
    const int64_t nloc = f.nown + f.nnext;
    for (int k = 0; k < MAX_NV; ++k) jv_base_[k] = 0;
    for (int k = 0; k < MAX_NVV; ++k) hvv_pos_[k] = 0;
    for (int k = 0; k < nv; ++k) jv_base_[k] = jac_.colptr[nvar_ - nv + k];
    for (int a = 0; a < nv; ++a)
        for (int b = 0; b <= a; ++b) {
            // -1: the entry is not in the pattern (detected patterns only keep structural nonzeros)
            hvv_pos_[a * (a + 1) / 2 + b] = hess_.find(nvar_ - nv + a + 1, nvar_ - nv + b + 1);
        }
    // tasks
    ntask_ = 0;
    auto add_task = [&](int step, int own) {
        task_[ntask_].step = step;
        task_[ntask_].own = own;
        task_[ntask_].map_off = ntask_ * lc.map_step;
        task_[ntask_].next_rows = 0;
        ++ntask_;
    };
    add_task(1, 1);
    if (N >= 3) add_task((int)N - 1, 0);
    if (N >= 2) add_task((int)N, 1);
    const bool step1_gives_next = (N <= 2);
    fp_map_off_ = ntask_ * lc.map_step;
    bc_map_off_ = fp_map_off_ + lc.map_fp;
    ip_map_off_ = bc_map_off_ + lc.map_bc;      // initial-path job (shifted path rows only), laid out like the final-path map
    edge_map_.assign(ip_map_off_ + (f.path_shift ? lc.map_fp : 0), -1);

    // edge position sets
    std::vector<int64_t> jpos, hpos, gpos;
    auto add_range = [](std::vector<int64_t>& v, int64_t a, int64_t b) { for (int64_t k = a; k < b; ++k) v.push_back(k); };
    const int64_t first_tail_col = std::max<int64_t>((N - 1) * B, B);  // 0-based column where blocks N, N+1 start
    add_range(jpos, jac_.colptr[0], jac_.colptr[std::min<int64_t>(B, nvar_ - nv)]);
    add_range(jpos, jac_.colptr[std::min<int64_t>(first_tail_col, nvar_ - nv)], jac_.colptr[nvar_ - nv]);
    add_range(hpos, hess_.colptr[0], hess_.colptr[std::min<int64_t>(B, nvar_ - nv)]);
    add_range(hpos, hess_.colptr[std::min<int64_t>(first_tail_col, nvar_ - nv)], hess_.colptr[nvar_ - nv]);
    add_range(gpos, 0, std::min<int64_t>(B, nvar_ - nv));
    add_range(gpos, std::min<int64_t>(first_tail_col, nvar_ - nv), nvar_ - nv);
    // Quantities no STEP contributes to are not summed over the grid (Cfg::RH / RG in ctd_kernels.cuh): the
    // (v,v) Hessian entries and the v entries of the gradient then belong to the edge blocks (the boundary and
    // final-path elements are the only writers; an entry nobody writes is a structural zero of the pattern)
    const bool vv_edge = f.nh_vv_emit == 0, gv_edge = f.g_v.n == 0;
    if (vv_edge)
        for (int k = 0; k < nv * (nv + 1) / 2; ++k) if (hvv_pos_[k] >= 0) hpos.push_back(hvv_pos_[k]);
    if (gv_edge) add_range(gpos, nvar_ - nv, nvar_);
    // v columns of the Jacobian: rows of the edge-owned steps, final path rows, boundary rows
    for (int k = 0; k < nv; ++k) {
        const int64_t col = nvar_ - nv + k;  // 0-based
        for (int64_t q = jac_.colptr[col]; q < jac_.colptr[col + 1]; ++q) {
            const int64_t r = jac_.rows[q] - 1;  // 0-based row
            bool edge_row = r >= N * CB;
            for (int t = 0; t < ntask_; ++t)
                if (task_[t].own && r >= (int64_t)(task_[t].step - 1) * CB && r < (int64_t)task_[t].step * CB) edge_row = true;
            if (edge_row) jpos.push_back(q);
        }
    }
    auto uniq = [](std::vector<int64_t>& v) { std::sort(v.begin(), v.end()); v.erase(std::unique(v.begin(), v.end()), v.end()); };
    uniq(jpos); uniq(hpos); uniq(gpos);
    ne_j_ = (int)jpos.size(); ne_h_ = (int)hpos.size(); ne_g_ = (int)gpos.size();
    edge_pos_.clear();
    edge_pos_.insert(edge_pos_.end(), jpos.begin(), jpos.end());
    edge_pos_.insert(edge_pos_.end(), hpos.begin(), hpos.end());
    edge_pos_.insert(edge_pos_.end(), gpos.begin(), gpos.end());
    auto slot_of = [](const std::vector<int64_t>& v, int64_t pos, int base) -> int {
        if (pos < 0) return -1;
        auto it = std::lower_bound(v.begin(), v.end(), pos);
        if (it == v.end() || *it != pos) throw CtdError{-9, "internal: edge value outside the edge blocks"};
        return base + (int)(it - v.begin());
    };
    // local variable a of step i -> 1-based global column
    auto gcol = [&](int64_t i, int a) -> int64_t {
        if (a < nloc) return (i - 1) * B + a + 1;
        return nvar_ - nv + (a - nloc) + 1;
    };
    // Cfg<E,F>::M_* offsets restated from the table sizes
    const int M_JO = 0, M_JN = M_JO + f.j_own.n, M_JV = M_JN + f.j_next.n, M_HO = M_JV + f.j_v.n,
              M_HN = M_HO + f.h_own.n, M_GO = M_HN + f.h_next.n, M_GN = M_GO + f.g_own.n, M_STEP = M_GN + f.g_next.n;
    if (M_STEP + 1 != lc.map_step) throw CtdError{-9, "internal: map layout mismatch"};
    for (int t = 0; t < ntask_; ++t) {
        int* m = edge_map_.data() + task_[t].map_off;
        const int64_t i = task_[t].step;
        const bool own = task_[t].own != 0;
        const bool next = (i >= N - 1) && (i != 1 || step1_gives_next || N == 1);
        const int64_t row0 = (i - 1) * CB;
        // Shifted rows (implicit Euler with path constraints): local rows >= eqs are the path rows of node i+1, i.e.
        // rows of the NEXT step's block; those of node N+1 belong to the final-path element (no entry: -1).
        auto shifted = [&](int r) { return f.path_shift && r >= f.eqs; };
        auto grow = [&](int r) -> int64_t { return shifted(r) ? (i < N ? i * CB + r + 1 : -1) : row0 + r + 1; };
        auto jfind = [&](int r, int64_t col) -> int64_t { const int64_t g = grow(r); return g < 0 ? -1 : jac_.find(g, col); };
        task_[t].next_rows = (f.path_shift && next && i < N) ? 1 : 0;
        m[M_STEP] = own ? 1 : 0;
        if (own) {
            for (int n = 0; n < f.j_own.n; ++n)
                m[M_JO + n] = slot_of(jpos, jfind(f.j_own(n, 0), gcol(i, f.j_own(n, 1))), 0);
        }
        // v-column entries: the step's own rows if it owns them, the shifted rows if it gives to the next block
        for (int n = 0; n < f.j_v.n; ++n) {
            const int r = f.j_v(n, 0);
            if (shifted(r) ? (next && i < N) : own)
                m[M_JV + n] = slot_of(jpos, jfind(r, nvar_ - nv + f.j_v(n, 1) + 1), 0);
        }
        if (own) {
            for (int n = 0; n < f.h_own.n; ++n) {
                const int64_t ga = gcol(i, f.h_own(n, 0)), gb = gcol(i, f.h_own(n, 1));
                m[M_HO + n] = slot_of(hpos, hess_.find(std::max(ga, gb), std::min(ga, gb)), ne_j_);
            }
            for (int n = 0; n < f.g_own.n; ++n) m[M_GO + n] = slot_of(gpos, gcol(i, f.g_own(n, 0)) - 1, ne_j_ + ne_h_);
        }
        if (next) {
            for (int n = 0; n < f.j_next.n; ++n)
                m[M_JN + n] = slot_of(jpos, jfind(f.j_next(n, 0), gcol(i, f.j_next(n, 1))), 0);
            for (int n = 0; n < f.h_next.n; ++n) {
                const int64_t ga = gcol(i, f.h_next(n, 0)), gb = gcol(i, f.h_next(n, 1));
                m[M_HN + n] = slot_of(hpos, hess_.find(std::max(ga, gb), std::min(ga, gb)), ne_j_);
            }
            for (int n = 0; n < f.g_next.n; ++n) m[M_GN + n] = slot_of(gpos, gcol(i, f.g_next(n, 0)) - 1, ne_j_ + ne_h_);
        }
    }
    // final path element: local (u_f block, x_f, v), or (x_f, u_f, v) for trapeze
    {
        int* m = edge_map_.data() + fp_map_off_;
        const int cbk = f.ctrl_blk;
        auto fcol = [&](int a) -> int64_t {
            const int nl = f.fp_nloc - f.nv;
            if (a >= nl) return nvar_ - nv + (a - nl) + 1;
            if (f.final_control) return N * B + a + 1;
            if (a < cbk) return (N - 1) * B + f.nx + a + 1;
            return N * B + (a - cbk) + 1;
        };
        for (int n = 0; n < f.fp_j.n; ++n)
            m[n] = slot_of(jpos, jac_.find(N * CB + f.fp_j(n, 0) + 1, fcol(f.fp_j(n, 1))), 0);
        for (int n = 0; n < f.fp_h.n; ++n) {
            const int a = f.fp_h(n, 0), b = f.fp_h(n, 1), nl = f.fp_nloc - f.nv;
            if (a >= nl && b >= nl) {
                const int ka = std::max(a, b) - nl, kb = std::min(a, b) - nl;
                const int kk = ka * (ka + 1) / 2 + kb;
                m[f.fp_j.n + n] = vv_edge ? slot_of(hpos, hvv_pos_[kk], ne_j_) : -2 - kk;
            } else {
                const int64_t ga = fcol(a), gb = fcol(b);
                m[f.fp_j.n + n] = slot_of(hpos, hess_.find(std::max(ga, gb), std::min(ga, gb)), ne_j_);
            }
        }
    }
    // initial-path job of shifted elements: the final-path code on (U_1 block, x_1, v), rows eqs.. of step 1
    if (f.path_shift) {
        int* m = edge_map_.data() + ip_map_off_;
        const int cbk = f.ctrl_blk, nl = f.fp_nloc - f.nv;
        auto icol = [&](int a) -> int64_t {
            if (a >= nl) return nvar_ - nv + (a - nl) + 1;
            if (a < cbk) return f.nx + a + 1;
            return (a - cbk) + 1;
        };
        for (int n = 0; n < f.fp_j.n; ++n)
            m[n] = slot_of(jpos, jac_.find(f.eqs + f.fp_j(n, 0) + 1, icol(f.fp_j(n, 1))), 0);
        for (int n = 0; n < f.fp_h.n; ++n) {
            const int a = f.fp_h(n, 0), b = f.fp_h(n, 1);
            if (a >= nl && b >= nl) {
                const int ka = std::max(a, b) - nl, kb = std::min(a, b) - nl;
                const int kk = ka * (ka + 1) / 2 + kb;
                m[f.fp_j.n + n] = vv_edge ? slot_of(hpos, hvv_pos_[kk], ne_j_) : -2 - kk;
            } else {
                const int64_t ga = icol(a), gb = icol(b);
                m[f.fp_j.n + n] = slot_of(hpos, hess_.find(std::max(ga, gb), std::min(ga, gb)), ne_j_);
            }
        }
    }
    // boundary + Mayer element: local (x_1, x_{N+1}, v)
    {
        int* m = edge_map_.data() + bc_map_off_;
        auto bcol = [&](int a) -> int64_t {
            if (a < f.nx) return a + 1;
            if (a < 2 * f.nx) return N * B + (a - f.nx) + 1;
            return nvar_ - nv + (a - 2 * f.nx) + 1;
        };
        for (int n = 0; n < f.bc_j.n; ++n)
            m[n] = slot_of(jpos, jac_.find(ncon_ - f.nb + f.bc_j(n, 0) + 1, bcol(f.bc_j(n, 1))), 0);
        for (int n = 0; n < f.bc_h.n; ++n) {
            const int a = f.bc_h(n, 0), b = f.bc_h(n, 1);
            if (a >= 2 * f.nx && b >= 2 * f.nx) {
                const int ka = std::max(a, b) - 2 * f.nx, kb = std::min(a, b) - 2 * f.nx;
                const int kk = ka * (ka + 1) / 2 + kb;
                m[f.bc_j.n + n] = vv_edge ? slot_of(hpos, hvv_pos_[kk], ne_j_) : -2 - kk;
            } else {
                const int64_t ga = bcol(a), gb = bcol(b);
                m[f.bc_j.n + n] = slot_of(hpos, hess_.find(std::max(ga, gb), std::min(ga, gb)), ne_j_);
            }
        }
        for (int n = 0; n < f.bc_g.n; ++n) {
            const int a = f.bc_g(n, 0);
            if (a >= 2 * f.nx) m[f.bc_j.n + f.bc_h.n + n] = gv_edge ? slot_of(gpos, nvar_ - nv + (a - 2 * f.nx), ne_j_ + ne_h_) : -2 - (a - 2 * f.nx);
            else m[f.bc_j.n + f.bc_h.n + n] = slot_of(gpos, bcol(a) - 1, ne_j_ + ne_h_);
        }
    }
}

// --------------------------------------------------------------------------------------------
void Model::ensure_device() {
    if (dev_ready_) {
        CTD_CUDA(cudaSetDevice(device_));
        return;
    }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        throw CtdError{(int)(e == cudaSuccess ? cudaErrorNoDevice : e),
                       "no CUDA device: this library has no CPU fallback (evaluation needs a B200)"};
    CTD_CUDA(cudaSetDevice(device_));
    CTD_CUDA(cudaStreamCreateWithFlags(&own_stream_, cudaStreamNonBlocking));
    if (!stream_) stream_ = own_stream_;
    const LaunchCfg lc = reg_.launcher.cfg(F_CONS | F_JAC | F_HESS | F_GRAD | F_OBJ);
    const int max_nred = lc.nred + 1;   // (+1: the violation slot of the flag sets that carry it)
    const int npar = std::max(reg_.info.npar, 1);
    CTD_CUDA(cudaMalloc(&d_edge_map_, sizeof(int) * std::max<size_t>(edge_map_.size(), 1)));
    CTD_CUDA(cudaMalloc(&d_edge_pos_, sizeof(long long) * std::max<size_t>(edge_pos_.size(), 1)));
    CTD_CUDA(cudaMalloc(&d_sgrid_, sizeof(double) * sgrid_.size()));
    CTD_CUDA(cudaMalloc(&d_params_, sizeof(double) * (size_t)batch_ * npar));
    int64_t max_ctas = 1;
    for (int fl : {1, 2, 4, 16, 17, 24, 7, 31, 49, 63}) {
        const LaunchCfg q = reg_.launcher.cfg(fl);
        if (q.threads == 0) continue;
        const int64_t tiles = (N_ >= 3) ? ((int64_t)N_ - 2 + q.own - 1) / q.own : 0;
        max_ctas = std::max<int64_t>(max_ctas, tiles + 1);   // (one warp per CTA is the smallest launch configuration)
        if (q.p_threads > 0) max_ctas = std::max<int64_t>(max_ctas, ((N_ >= 3) ? ((int64_t)N_ - 2 + q.p_own - 1) / q.p_own : 0) + 1);
    }
    CTD_CUDA(cudaMalloc(&d_partials_, sizeof(double) * (size_t)batch_ * max_ctas * max_nred));
    CTD_CUDA(cudaMalloc(&d_counters_, sizeof(unsigned int) * batch_));
    CTD_CUDA(cudaMemsetAsync(d_counters_, 0, sizeof(unsigned int) * batch_, stream_));
    CTD_CUDA(cudaMemcpyAsync(d_edge_map_, edge_map_.data(), sizeof(int) * edge_map_.size(), cudaMemcpyHostToDevice, stream_));
    static_assert(sizeof(long long) == sizeof(int64_t), "int64 layout");
    CTD_CUDA(cudaMemcpyAsync(d_edge_pos_, edge_pos_.data(), sizeof(long long) * edge_pos_.size(), cudaMemcpyHostToDevice, stream_));
    CTD_CUDA(cudaMemcpyAsync(d_sgrid_, sgrid_.data(), sizeof(double) * sgrid_.size(), cudaMemcpyHostToDevice, stream_));
    CTD_CUDA(cudaMemcpyAsync(d_params_, params_.data(), sizeof(double) * params_.size(), cudaMemcpyHostToDevice, stream_));
    CTD_CUDA(cudaStreamSynchronize(stream_));
    dev_ready_ = true;
}

void Model::ensure_staging() {
    if (staging_ready_) return;
    const size_t b = (size_t)batch_;
    CTD_CUDA(cudaMalloc(&d_x_, sizeof(double) * b * nvar_));
    CTD_CUDA(cudaMalloc(&d_y_, sizeof(double) * b * ncon_));
    CTD_CUDA(cudaMalloc(&d_c_, sizeof(double) * b * ncon_));
    CTD_CUDA(cudaMalloc(&d_j_, sizeof(double) * b * std::max<int64_t>(nnzj(), 1)));
    CTD_CUDA(cudaMalloc(&d_h_, sizeof(double) * b * std::max<int64_t>(nnzh(), 1)));
    CTD_CUDA(cudaMalloc(&d_g_, sizeof(double) * b * nvar_));
    CTD_CUDA(cudaMalloc(&d_obj_, sizeof(double) * b));
    staging_ready_ = true;
}

void Model::set_params(const double* params, int n_params) {
    const ElemInfo& f = reg_.info;
    const int npar = std::max(f.npar, 1);
    if (n_params != f.npar && n_params != f.npar * batch_) throw CtdError{-2, "params must hold npar or batch*npar values"};
    for (int b = 0; b < batch_; ++b)
        for (int k = 0; k < f.npar; ++k)
            params_[(size_t)b * npar + k] = params[n_params == f.npar ? k : (size_t)b * f.npar + k];
    if (dev_ready_) {
        CTD_CUDA(cudaSetDevice(device_));
        CTD_CUDA(cudaMemcpyAsync(d_params_, params_.data(), sizeof(double) * params_.size(), cudaMemcpyHostToDevice, stream_));
        CTD_CUDA(cudaStreamSynchronize(stream_));
    }
}

cudaStream_t Model::stream() {
    ensure_device();
    return stream_;
}

void Model::set_stream(cudaStream_t s) {
    ensure_device();
    if ((s ? s : own_stream_) == stream_) return;   // already there: nothing to order
    // work already enqueued on the previous stream is not ordered with the new one: finish it first
    CTD_CUDA(cudaSetDevice(device_));
    CTD_CUDA(cudaStreamSynchronize(stream_));
    stream_ = s ? s : own_stream_;
}

void Model::sync() {
    if (dev_ready_) {
        CTD_CUDA(cudaSetDevice(device_));
        CTD_CUDA(cudaStreamSynchronize(stream_));
        if (d_peer_err_) {
            int e = 0;
            CTD_CUDA(cudaMemcpy(&e, d_peer_err_, sizeof(int), cudaMemcpyDeviceToHost));
            if (e) {
                cudaMemset(d_peer_err_, 0, sizeof(int));
                throw CtdError{-8, "ctd_peer_wait: rank " + std::to_string(e - 1) + " did not deliver its norms within 5 s"};
            }
        }
    }
}

EvalArgs Model::base_args() const {
    EvalArgs a;
    std::memset(&a, 0, sizeof(a));
    a.params = d_params_;
    a.x_stride = nvar_; a.y_stride = ncon_; a.c_stride = ncon_; a.j_stride = nnzj(); a.h_stride = nnzh(); a.g_stride = nvar_;
    a.sgrid = d_sgrid_;
    a.N = N_;
    a.nvar = nvar_; a.ncon = ncon_;
    for (int k = 0; k < MAX_NV; ++k) a.jv_base[k] = jv_base_[k];
    for (int k = 0; k < MAX_NVV; ++k) a.hvv_pos[k] = hvv_pos_[k];
    a.partials = d_partials_;
    a.counters = d_counters_;
    a.edge_map = d_edge_map_;
    a.edge_pos = d_edge_pos_;
    a.ne_j = ne_j_; a.ne_h = ne_h_; a.ne_g = ne_g_;
    a.ntask = ntask_;
    for (int t = 0; t < MAX_TASKS; ++t) a.task[t] = task_[t];
    a.fp_map_off = fp_map_off_; a.bc_map_off = bc_map_off_; a.ip_map_off = ip_map_off_;
    a.nmap = (int)edge_map_.size();
    return a;
}

void Model::eval_dev(int flags, const double* x, const double* y, double sigma, double* obj, double* grad, double* c,
                     double* jvals, double* hvals, double* viol) {
    ensure_device();
    flags &= ~F_VIOL;
    if (viol) {
        if (!(flags & F_CONS)) throw CtdError{-2, "the violation norm needs the constraints (CTD_CONS)"};
        if (reg_.launcher.cfg(flags | F_VIOL).threads > 0) {
            flags |= F_VIOL;          // reduced inside the evaluation kernel
        } else {                      // wide elements: evaluation, then the stand-alone norm kernel
            eval_dev(flags, x, y, sigma, obj, grad, c, jvals, hvals, nullptr);
            violation_dev(c, viol);
            return;
        }
    }
    const LaunchCfg lc = reg_.launcher.cfg(flags);
    if (lc.threads == 0) throw CtdError{-2, "unsupported combination of outputs"};
    if (!x) throw CtdError{-2, "x is NULL"};
    if ((flags & F_CONS) && !c) throw CtdError{-2, "c is NULL"};
    if ((flags & F_JAC) && !jvals) throw CtdError{-2, "jvals is NULL"};
    if ((flags & F_HESS) && (!hvals || !y)) throw CtdError{-2, "hvals / y is NULL"};
    if ((flags & F_GRAD) && !grad) throw CtdError{-2, "grad is NULL"};
    EvalArgs a = base_args();
    a.x = x; a.y = y; a.sigma = sigma; a.obj = obj; a.grad = grad; a.c = c; a.jvals = jvals; a.hvals = hvals; a.viol = viol;
    if (peer_push_) {   // ctd_eval_gather_dev: the reduction epilogue also pushes the norms to every rank
        if (!(flags & F_VIOL)) throw CtdError{-6, "the fused gather needs the in-kernel violation norm (light elements)"};
        peer_args(a);
    }
    // warps per CTA: large CTAs for a single instance (a 1e5-step grid is one wave and CTA launch rate
    // matters), small CTAs for batches (many waves: resources are released per CTA, and the tiles of
    // an instance rarely fill a large CTA)
    // (measured: N = 1000 x 4096 instances 78 % -> 94 % of the HBM peak with 3 warps per CTA, whereas a
    // single 1e6-step grid is faster with 8)
    static const int wpc_env = getenv("CTD_WPC") ? atoi(getenv("CTD_WPC")) : 0;   // tuning runs only
    a.tiles = (N_ >= 3) ? (int)(((int64_t)N_ - 2 + lc.own - 1) / lc.own) : 0;
    // A single instance whose tiles are all resident at once (one wave: <= 24 warps on each of the 148 SMs) is
    // launched as 2-warp CTAs: 3226 tiles of a 1e5-step grid then spread as 22 warps on EVERY SM instead of a mix of
    // 24 and 16 (8-warp CTAs), and with programmatic dependent launch the CTAs are scheduled while the previous
    // kernel drains, so their number costs nothing (10.30 -> 9.95 us per evaluation, one stream).
    int wpc = batch_ > 1 ? std::min(lc.wpc, 3) : lc.wpc;
    if (batch_ == 1 && lc.wpc >= 8 && a.tiles <= 148 * 24) wpc = 2;
    if (wpc_env > 0) wpc = std::min(lc.wpc, wpc_env);
    int threads = 32 * wpc;
    a.ctas = (a.tiles + wpc - 1) / wpc + 1;
    // wide elements on small grids: one tile per CTA, the tile's outputs split over the CTA's warps
    // (enough warps to occupy the GPU although the grid has few steps)
    int parts = 0;
    size_t tile_smem = wpc * lc.warp_smem + sizeof(double) * (wpc * (lc.nred + 1) + 2);
    if (lc.p_threads > 0) {
        const int64_t ptiles = (N_ >= 3) ? ((int64_t)N_ - 2 + lc.p_own - 1) / lc.p_own : 0;
        // (measured: robot/trapeze with 3226 tiles is 40 % faster through the one-warp-per-tile kernel, the
        //  3-stage IRK at 645 tiles 2.5x faster through the parts kernel)
        bool use_parts = ptiles * batch_ * (lc.p_threads / 32) <= (int64_t)148 * 16;
#ifdef CTD_TRACE
        if (getenv("CTD_NO_PARTS")) use_parts = false;   // debug builds: the timeline exists only in the one-warp kernel
#endif
        if (use_parts) {
            parts = 1;
            threads = lc.p_threads;
            a.tiles = (int)ptiles;
            a.ctas = (int)ptiles + 1;
            tile_smem = lc.p_smem;
        }
    }
    // the edge CTA uses the same dynamic shared memory as its value buffer (its reduction scratch at
    // S_RED is only touched after the buffer has been scattered)
    // (layers, position maps, one staging block per step task, destination positions)
    const size_t ne_all = (size_t)ne_j_ + ne_h_ + ne_g_;
    const size_t edge_smem = sizeof(double) * (edge_layers(reg_.info.path_shift) * ne_all + 2) + sizeof(int) * (edge_map_.size() + 2) +
                             sizeof(double) * (size_t)ntask_ * (size_t)lc.map_step + sizeof(long long) * (ne_all + 1);
    size_t smem = std::max(tile_smem, edge_smem);
    // (tuning runs only: a shared-memory floor caps the CTAs resident per SM.  Capping a one-wave launch at
    //  ceil(ctas / 148) per SM to spread it evenly was measured and NOT adopted: it helps when the inputs hit in L2
    //  (4 rotating buffer sets: 9.02 -> 8.85 us) and hurts when they do not (6 sets: 9.72 -> 10.04, 8 sets: 9.94 -> 10.43;
    //  bench.py: 9.54 -> 9.63), see DESIGN.md.)
    static const size_t smem_min_env = getenv("CTD_SMEM_MIN") ? (size_t)atoll(getenv("CTD_SMEM_MIN")) : 0;
    if (smem_min_env > smem && batch_ == 1) smem = smem_min_env;
    // The dynamic shared-memory opt-in is an attribute of the KERNEL (per device, for the whole process), not
    // of this handle: another handle on the same element may have raised it already, and lowering it would
    // break that handle's launches.  It only ever grows, under a process-wide lock.
    {
        static std::mutex mu;
        static std::map<std::tuple<int, const void*, int>, size_t> optin;   // (device, element, flag set) -> bytes
        std::lock_guard<std::mutex> lock(mu);
        size_t& cur = optin[std::make_tuple(device_, (const void*)reg_.launcher.launch, flags + 256 * parts)];
        if (cur < smem) {
            const int e = reg_.launcher.prepare(flags, smem, parts);
            if (e != 0) throw CtdError{e, std::string("cudaFuncSetAttribute failed: ") + cudaGetErrorString((cudaError_t)e)};
            cur = smem;
        }
    }
#ifdef CTD_TRACE
    // debug builds only: per-warp timestamps, dumped as raw int64 to $CTD_TRACE_FILE.  Without CTD_TRACE_AT every
    // launch is traced on its own (memset before, copy + synchronize after).  With CTD_TRACE_AT=K the launches K-1
    // and K of this handle are traced into one preallocated buffer with nothing but the two kernels on the stream
    // between them: the timeline of a launch in the steady state of back-to-back (PDL) launches.
    long long* d_trace = nullptr;
    const size_t ntrace = (size_t)a.ctas * std::max(lc.wpc, std::max(threads, lc.p_threads) / 32) * 8;
    const char* trace_file = getenv("CTD_TRACE_FILE");
    const char* trace_at = getenv("CTD_TRACE_AT");
    static long long* d_pair = nullptr;
    const int64_t at = trace_at ? atoll(trace_at) : -1;
    if (trace_file && at >= 0) {
        if (!d_pair) {
            CTD_CUDA(cudaMalloc(&d_pair, 2 * ntrace * sizeof(long long)));
            CTD_CUDA(cudaMemsetAsync(d_pair, 0, 2 * ntrace * sizeof(long long), stream_));
            CTD_CUDA(cudaStreamSynchronize(stream_));
        }
        if (launches_ == at - 1) a.trace = d_pair;
        if (launches_ == at) a.trace = d_pair + ntrace;
    } else if (trace_file) {
        CTD_CUDA(cudaMalloc(&d_trace, ntrace * sizeof(long long)));
        CTD_CUDA(cudaMemsetAsync(d_trace, 0, ntrace * sizeof(long long), stream_));
        a.trace = d_trace;
    }
#endif
    const int e = reg_.launcher.launch(a, flags, batch_, smem, stream_, parts, threads);
    if (e != 0) throw CtdError{e, std::string("kernel launch failed: ") + cudaGetErrorString((cudaError_t)e)};
    ++launches_;
#ifdef CTD_TRACE
    if (d_trace) {
        std::vector<long long> h(ntrace);
        CTD_CUDA(cudaMemcpyAsync(h.data(), d_trace, ntrace * sizeof(long long), cudaMemcpyDeviceToHost, stream_));
        CTD_CUDA(cudaStreamSynchronize(stream_));
        if (FILE* f = fopen(trace_file, "wb")) { fwrite(h.data(), sizeof(long long), ntrace, f); fclose(f); }
        cudaFree(d_trace);
    }
    if (d_pair && launches_ == at + 8) {   // a few launches later: the traced pair had successors like any other
        std::vector<long long> h(2 * ntrace);
        CTD_CUDA(cudaMemcpyAsync(h.data(), d_pair, 2 * ntrace * sizeof(long long), cudaMemcpyDeviceToHost, stream_));
        CTD_CUDA(cudaStreamSynchronize(stream_));
        if (FILE* f = fopen(trace_file, "wb")) { fwrite(h.data(), sizeof(long long), 2 * ntrace, f); fclose(f); }
    }
#endif
}

void Model::eval_host(int flags, const double* x, const double* y, double sigma, double* obj, double* grad, double* c,
                      double* jvals, double* hvals, bool wait) {
    ensure_device();
    ensure_staging();
    const size_t b = (size_t)batch_;
    if (!x) throw CtdError{-2, "x is NULL"};
    CTD_CUDA(cudaMemcpyAsync(d_x_, x, sizeof(double) * b * nvar_, cudaMemcpyHostToDevice, stream_));
    if (flags & F_HESS) {
        if (!y) throw CtdError{-2, "y is NULL"};
        CTD_CUDA(cudaMemcpyAsync(d_y_, y, sizeof(double) * b * ncon_, cudaMemcpyHostToDevice, stream_));
    }
    eval_dev(flags, d_x_, d_y_, sigma, d_obj_, d_g_, d_c_, d_j_, d_h_);
    if ((flags & F_CONS) && c) CTD_CUDA(cudaMemcpyAsync(c, d_c_, sizeof(double) * b * ncon_, cudaMemcpyDeviceToHost, stream_));
    if ((flags & F_JAC) && jvals) CTD_CUDA(cudaMemcpyAsync(jvals, d_j_, sizeof(double) * b * nnzj(), cudaMemcpyDeviceToHost, stream_));
    if ((flags & F_HESS) && hvals) CTD_CUDA(cudaMemcpyAsync(hvals, d_h_, sizeof(double) * b * nnzh(), cudaMemcpyDeviceToHost, stream_));
    if ((flags & F_GRAD) && grad) CTD_CUDA(cudaMemcpyAsync(grad, d_g_, sizeof(double) * b * nvar_, cudaMemcpyDeviceToHost, stream_));
    if ((flags & F_OBJ) && obj) CTD_CUDA(cudaMemcpyAsync(obj, d_obj_, sizeof(double) * b, cudaMemcpyDeviceToHost, stream_));
    if (wait) CTD_CUDA(cudaStreamSynchronize(stream_));
}

// --------------------------------------------------------------------------------------------
// Per-instance constraint violation  max_k max(lcon_k - c_k, c_k - ucon_k, 0): one CTA per instance,
// coalesced pass over c, warp-shuffle + shared-memory max reduction.
__global__ void __launch_bounds__(256) ctd_violation_kernel(const double* __restrict__ c, const double* __restrict__ lcon,
                                                            const double* __restrict__ ucon, long long ncon,
                                                            double* __restrict__ viol) {
    const double* ci = c + (long long)blockIdx.x * ncon;
    double m = 0.0;
    for (long long k = threadIdx.x; k < ncon; k += blockDim.x) {
        const double v = ci[k];
        const double d = fmax(lcon[k] - v, v - ucon[k]);
        m = (d > m || d != d) ? d : m;  // NaN propagates
    }
    __shared__ double sm[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double w = __shfl_xor_sync(0xffffffffu, m, o);
        m = (w > m || w != w) ? w : m;
    }
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) m = (sm[w] > m || sm[w] != sm[w]) ? sm[w] : m;
        viol[blockIdx.x] = m;
    }
}

void Model::unpack_dev(const double* x, const double* y, double* time, double* state, double* control, double* variable,
                       double* costate, double* path_dual, double* boundary_dual) {
    ensure_device();
    if (!x) throw CtdError{-2, "x is NULL"};
    UnpackArgs a;
    a.x = x; a.y = y; a.time = time; a.state = state; a.control = control; a.variable = variable;
    a.costate = costate; a.path_dual = path_dual; a.boundary_dual = boundary_dual;
    a.sgrid = d_sgrid_; a.params = d_params_; a.nvar = nvar_; a.ncon = ncon_; a.N = N_;
    const int e = reg_.launcher.unpack(a, batch_, stream_);
    if (e != 0) throw CtdError{e, std::string("kernel launch failed: ") + cudaGetErrorString((cudaError_t)e)};
    ++launches_;
}

void Model::unpack_host(const double* x, const double* y, double* time, double* state, double* control, double* variable,
                        double* costate, double* path_dual, double* boundary_dual) {
    ensure_device();
    ensure_staging();
    const ElemInfo& f = reg_.info;
    const size_t b = (size_t)batch_, n1 = (size_t)N_ + 1;
    const size_t nc = (size_t)N_ * f.control_steps + 1;   // rows of the control array (common.jl:82-99)
    const size_t sz[7] = {b * n1, b * n1 * f.nx, b * nc * f.nu, b * f.nv, b * N_ * f.nx, b * n1 * f.np, b * f.nb};
    double* host[7] = {time, state, control, variable, costate, path_dual, boundary_dual};
    double* dev[7] = {nullptr};
    size_t total = 0;
    for (int k = 0; k < 7; ++k) if (host[k] && sz[k]) total += sz[k];
    double* pool = nullptr;
    if (total) CTD_CUDA(cudaMalloc(&pool, sizeof(double) * total));
    size_t off = 0;
    for (int k = 0; k < 7; ++k) if (host[k] && sz[k]) { dev[k] = pool + off; off += sz[k]; }
    try {
        CTD_CUDA(cudaMemcpyAsync(d_x_, x, sizeof(double) * b * nvar_, cudaMemcpyHostToDevice, stream_));
        if (y) CTD_CUDA(cudaMemcpyAsync(d_y_, y, sizeof(double) * b * ncon_, cudaMemcpyHostToDevice, stream_));
        unpack_dev(d_x_, y ? d_y_ : nullptr, dev[0], dev[1], dev[2], dev[3], dev[4], dev[5], dev[6]);
        for (int k = 0; k < 7; ++k)
            if (dev[k]) CTD_CUDA(cudaMemcpyAsync(host[k], dev[k], sizeof(double) * sz[k], cudaMemcpyDeviceToHost, stream_));
        CTD_CUDA(cudaStreamSynchronize(stream_));
    } catch (...) {
        cudaFree(pool);
        throw;
    }
    cudaFree(pool);
}

void Model::pack_dev(const double* state, const double* control, const double* variable, const double* stage_control,
                     double* x) {
    ensure_device();
    if (!x) throw CtdError{-2, "x is NULL"};
    PackArgs a;
    a.state = state; a.control = control; a.variable = variable; a.stage_control = stage_control;
    a.x = x; a.nvar = nvar_; a.N = N_;
    const int e = reg_.launcher.pack(a, batch_, stream_);
    if (e != 0) throw CtdError{e, std::string("kernel launch failed: ") + cudaGetErrorString((cudaError_t)e)};
    ++launches_;
}

void Model::pack_host(const double* state, const double* control, const double* variable, const double* stage_control,
                      double* x) {
    ensure_device();
    ensure_staging();
    const ElemInfo& f = reg_.info;
    const size_t b = (size_t)batch_, n1 = (size_t)N_ + 1;
    const size_t sz[4] = {b * n1 * f.nx, b * n1 * f.nu, b * f.nv, b * N_ * f.stage * f.nu};
    const double* host[4] = {state, control, variable, stage_control};
    double* dev[4] = {nullptr, nullptr, nullptr, nullptr};
    size_t total = 0;
    for (int k = 0; k < 4; ++k) if (host[k] && sz[k]) total += sz[k];
    double* pool = nullptr;
    if (total) CTD_CUDA(cudaMalloc(&pool, sizeof(double) * total));
    size_t off = 0;
    try {
        for (int k = 0; k < 4; ++k)
            if (host[k] && sz[k]) {
                dev[k] = pool + off;
                off += sz[k];
                CTD_CUDA(cudaMemcpyAsync(dev[k], host[k], sizeof(double) * sz[k], cudaMemcpyHostToDevice, stream_));
            }
        pack_dev(dev[0], dev[1], dev[2], dev[3], d_x_);
        CTD_CUDA(cudaMemcpyAsync(x, d_x_, sizeof(double) * b * nvar_, cudaMemcpyDeviceToHost, stream_));
        CTD_CUDA(cudaStreamSynchronize(stream_));
    } catch (...) {
        cudaFree(pool);
        throw;
    }
    cudaFree(pool);
}

void Model::ensure_con_bounds() {
    if (d_lcon_) return;
    std::vector<double> l(ncon_), u(ncon_);
    constraint_bounds(l.data(), u.data());
    CTD_CUDA(cudaMalloc(&d_lcon_, sizeof(double) * ncon_));
    CTD_CUDA(cudaMalloc(&d_ucon_, sizeof(double) * ncon_));
    CTD_CUDA(cudaMalloc(&d_viol_, sizeof(double) * batch_));
    CTD_CUDA(cudaMemcpyAsync(d_lcon_, l.data(), sizeof(double) * ncon_, cudaMemcpyHostToDevice, stream_));
    CTD_CUDA(cudaMemcpyAsync(d_ucon_, u.data(), sizeof(double) * ncon_, cudaMemcpyHostToDevice, stream_));
    CTD_CUDA(cudaStreamSynchronize(stream_));
}

void Model::violation_dev(const double* c, double* viol) {
    ensure_device();
    ensure_con_bounds();
    if (!c || !viol) throw CtdError{-2, "c / viol is NULL"};
    ctd_violation_kernel<<<batch_, 256, 0, stream_>>>(c, d_lcon_, d_ucon_, ncon_, viol);
    CTD_CUDA(cudaGetLastError());
    ++launches_;
}

void Model::norms_dev(const double* x, double* obj, double* viol) {
    ensure_device();
    ensure_staging();
    ensure_con_bounds();
    if (!viol) viol = d_viol_;
    if (!obj) obj = d_obj_;
    // objective and violation norm come out of ONE launch (the violation is a max-reduction slot of the
    // fused kernel); the constraint values are a by-product in the handle's buffer
    eval_dev(F_CONS | F_OBJ, x, nullptr, 1.0, obj, nullptr, d_c_, nullptr, nullptr, viol);
}

void Model::norms_host(const double* x, double* obj, double* viol) {
    ensure_device();
    ensure_staging();
    const size_t b = (size_t)batch_;
    CTD_CUDA(cudaMemcpyAsync(d_x_, x, sizeof(double) * b * nvar_, cudaMemcpyHostToDevice, stream_));
    norms_dev(d_x_, nullptr, nullptr);
    CTD_CUDA(cudaMemcpyAsync(obj, d_obj_, sizeof(double) * b, cudaMemcpyDeviceToHost, stream_));
    CTD_CUDA(cudaMemcpyAsync(viol, d_viol_, sizeof(double) * b, cudaMemcpyDeviceToHost, stream_));
    CTD_CUDA(cudaStreamSynchronize(stream_));
}

}  // namespace ctd
