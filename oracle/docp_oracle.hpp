// TEST INFRASTRUCTURE ONLY -- CPU oracle for the DOCP evaluation path.
//
// Restatement, in plain C++17, of the reference's (CTDirect.jl v1.0.12)
// discretised-OCP functions.  Nothing here is shipped or timed as the product;
// only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference
// arm may call it.
//
// PARITY STATUS: structure counts, dimensions and the known-answer constraint /
// objective values are pinned against the reference's own tests
// (test/ci/test_modeler_solver.jl:37, test/ci/test_discretization_stagewise.jl:93-99,
// 176-198, test/ci/test_zero_control_allocations.jl:31,138, test/archives/AD_backend.md:57-64).
// Jacobian / Hessian VALUES and the COO ORDERING are "parity unpinned": the
// reference stores no golden vectors for them and its arithmetic lives in the
// un-vendored ADNLPModels 0.8 / ForwardDiff / ReverseDiff / SparseArrays
// packages; Julia is not available in this image.  Ordering follows the
// documented conventions (CSC sort of SparseArrays.sparse; tril for the Hessian).
//
// Each function cites the reference file:line it follows (paths relative to
// /root/reference).  Indices in comments are 1-based as in Julia; C++ storage is
// 0-based.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "hdual.hpp"
#include "tracer.hpp"
#include "problems.hpp"

namespace oracle {

// scheme ids shared with the C API (oracle_capi.cpp)
enum SchemeId {
    TRAPEZE = 0,          // src/ode/trapeze.jl
    MIDPOINT = 1,         // src/ode/midpoint.jl
    GL1_CONST = 2,        // src/ode/irk.jl:20-51 (test only)
    GL2_CONST = 3,        // :gauss_legendre_2_constant_control  irk.jl:58-89
    GL3_CONST = 4,        // :gauss_legendre_3_constant_control  irk.jl:96-131
    GL2_STAGEWISE = 5,    // :gauss_legendre_2  irk_stagewise.jl
    GL3_STAGEWISE = 6,    // :gauss_legendre_3  irk_stagewise.jl
    EULER_EXPLICIT = 7,   // :euler | :euler_explicit | :euler_forward   src/ode/euler.jl
    EULER_IMPLICIT = 8,   // :euler_implicit | :euler_backward           src/ode/euler.jl
    // midpoint with several controls per time step: the DirectShooting discretizer (src/direct_shooting.jl:55-71 ->
    // DOCP(ocp, grid_size, control_steps, :midpoint, nothing); midpoint.jl:20,57-65,98-113,145-154)
    MIDPOINT_CS2 = 9,
    MIDPOINT_CS3 = 10,
    MIDPOINT_CS10 = 11,
};

inline int scheme_from_name(const std::string& s) {
    // symbol switch of src/DOCP_data.jl:307-349
    if (s == "trapeze") return TRAPEZE;
    if (s == "midpoint") return MIDPOINT;
    if (s == "gauss_legendre_1_constant_control") return GL1_CONST;
    if (s == "gauss_legendre_2_constant_control") return GL2_CONST;
    if (s == "gauss_legendre_3_constant_control") return GL3_CONST;
    if (s == "gauss_legendre_2") return GL2_STAGEWISE;
    if (s == "gauss_legendre_3") return GL3_STAGEWISE;
    if (s == "euler" || s == "euler_explicit" || s == "euler_forward") return EULER_EXPLICIT;   // DOCP_data.jl:315
    if (s == "euler_implicit" || s == "euler_backward") return EULER_IMPLICIT;                  // DOCP_data.jl:318
    if (s == "midpoint_cs2") return MIDPOINT_CS2;
    if (s == "midpoint_cs3") return MIDPOINT_CS3;
    if (s == "midpoint_cs10") return MIDPOINT_CS10;
    return -1;
}

template <int SCH> struct SchemeTraits;
template <> struct SchemeTraits<TRAPEZE> { static constexpr int stage = 0; static constexpr bool irk = false, stagewise = false, final_control = true; static constexpr int control_steps = 1; static constexpr bool midpoint = false; };
template <> struct SchemeTraits<MIDPOINT> { static constexpr int stage = 0; static constexpr bool irk = false, stagewise = false, final_control = false; static constexpr int control_steps = 1; static constexpr bool midpoint = true; };
template <> struct SchemeTraits<GL1_CONST> { static constexpr int stage = 1; static constexpr bool irk = true, stagewise = false, final_control = false; static constexpr int control_steps = 1; static constexpr bool midpoint = false; };
template <> struct SchemeTraits<GL2_CONST> { static constexpr int stage = 2; static constexpr bool irk = true, stagewise = false, final_control = false; static constexpr int control_steps = 1; static constexpr bool midpoint = false; };
template <> struct SchemeTraits<GL3_CONST> { static constexpr int stage = 3; static constexpr bool irk = true, stagewise = false, final_control = false; static constexpr int control_steps = 1; static constexpr bool midpoint = false; };
template <> struct SchemeTraits<GL2_STAGEWISE> { static constexpr int stage = 2; static constexpr bool irk = true, stagewise = true, final_control = false; static constexpr int control_steps = 1; static constexpr bool midpoint = false; };
template <> struct SchemeTraits<GL3_STAGEWISE> { static constexpr int stage = 3; static constexpr bool irk = true, stagewise = true, final_control = false; static constexpr int control_steps = 1; static constexpr bool midpoint = false; };
template <> struct SchemeTraits<EULER_EXPLICIT> { static constexpr int stage = 0; static constexpr bool irk = false, stagewise = false, final_control = false; static constexpr int control_steps = 1; static constexpr bool midpoint = false; };
template <int CS> struct MidpointControlSteps { static constexpr int stage = 0; static constexpr bool irk = false, stagewise = false, final_control = false; static constexpr int control_steps = CS; static constexpr bool midpoint = true; };
template <> struct SchemeTraits<MIDPOINT_CS2> : MidpointControlSteps<2> {};
template <> struct SchemeTraits<MIDPOINT_CS3> : MidpointControlSteps<3> {};
template <> struct SchemeTraits<MIDPOINT_CS10> : MidpointControlSteps<10> {};
template <> struct SchemeTraits<EULER_IMPLICIT> { static constexpr int stage = 0; static constexpr bool irk = false, stagewise = false, final_control = false; static constexpr int control_steps = 1; static constexpr bool midpoint = false; };

// Butcher tableaux, computed from the same expressions as irk.jl:41-43,78-80,116-122
struct Butcher {
    int s = 0;
    double a[3][3] = {{0}};
    double b[3] = {0};
    double c[3] = {0};
};
inline Butcher butcher_for_stage(int s) {
    Butcher t;
    t.s = s;
    if (s == 1) {
        t.a[0][0] = 0.5; t.b[0] = 1; t.c[0] = 0.5;
    } else if (s == 2) {
        t.a[0][0] = 0.25; t.a[0][1] = 0.25 - std::sqrt(3.0) / 6;
        t.a[1][0] = 0.25 + std::sqrt(3.0) / 6; t.a[1][1] = 0.25;
        t.b[0] = 0.5; t.b[1] = 0.5;
        t.c[0] = 0.5 - std::sqrt(3.0) / 6; t.c[1] = 0.5 + std::sqrt(3.0) / 6;
    } else if (s == 3) {
        t.a[0][0] = 5.0 / 36.0; t.a[0][1] = 2.0 / 9 - std::sqrt(15.0) / 15; t.a[0][2] = 5.0 / 36 - std::sqrt(15.0) / 30;
        t.a[1][0] = 5.0 / 36.0 + std::sqrt(15.0) / 24; t.a[1][1] = 2.0 / 9.0; t.a[1][2] = 5.0 / 36.0 - std::sqrt(15.0) / 24;
        t.a[2][0] = 5.0 / 36 + std::sqrt(15.0) / 30; t.a[2][1] = 2.0 / 9 + std::sqrt(15.0) / 15; t.a[2][2] = 5.0 / 36.0;
        t.b[0] = 5.0 / 18.0; t.b[1] = 4.0 / 9.0; t.b[2] = 5.0 / 18.0;
        t.c[0] = 0.5 - 0.1 * std::sqrt(15.0); t.c[1] = 0.5; t.c[2] = 0.5 + 0.1 * std::sqrt(15.0);
    }
    return t;
}

// ---------------------------------------------------------------------------
// sparse pattern helpers: add_nonzero_block! (ode/common.jl:295-312) followed by
// SparseArrays.sparse(Is, Js, trues, m, n): CSC, sorted by column then row,
// duplicates merged.
// ---------------------------------------------------------------------------
struct Pattern {
    int64_t m = 0, n = 0;
    std::vector<int64_t> colptr;  // n+1, 0-based offsets
    std::vector<int64_t> rows;    // 1-based row indices, CSC order
    std::vector<int64_t> cols;    // 1-based col indices, CSC order
    int64_t nnz() const { return (int64_t)rows.size(); }
    // position of (r,c) (1-based) or -1
    int64_t find(int64_t r, int64_t c) const {
        if (c < 1 || c > n) return -1;
        auto b = rows.begin() + colptr[c - 1], e = rows.begin() + colptr[c];
        auto it = std::lower_bound(b, e, r);
        if (it == e || *it != r) return -1;
        return it - rows.begin();
    }
};

struct TripletList {
    std::vector<std::pair<int64_t, int64_t>> cr;  // (col,row) so that sort gives CSC order
    // ranges are inclusive, 1-based; empty when end<start (Julia range semantics)
    void add_block(int64_t i0, int64_t i1, int64_t j0, int64_t j1, bool sym = false) {
        for (int64_t i = i0; i <= i1; ++i)
            for (int64_t j = j0; j <= j1; ++j) {
                cr.emplace_back(j, i);
                if (sym) cr.emplace_back(i, j);
            }
    }
    void add(int64_t i, int64_t j) { cr.emplace_back(j, i); }
    Pattern build(int64_t m, int64_t n, bool lower_only) {
        std::sort(cr.begin(), cr.end());
        cr.erase(std::unique(cr.begin(), cr.end()), cr.end());
        Pattern p;
        p.m = m; p.n = n;
        p.colptr.assign(n + 1, 0);
        for (auto& e : cr) {
            if (lower_only && e.second < e.first) continue;  // tril keeps row >= col
            p.rows.push_back(e.second);
            p.cols.push_back(e.first);
            p.colptr[e.first]++;
        }
        for (int64_t j = 0; j < n; ++j) p.colptr[j + 1] += p.colptr[j];
        return p;
    }
};

// ---------------------------------------------------------------------------
// accessors: the restated functions read NLP variables through an accessor so
// the same code runs on plain doubles and on locally-seeded hyper-duals.
// ---------------------------------------------------------------------------
struct PlainAcc {
    using T = double;
    const double* x;
    double operator()(int j) const { return x[j]; }
};

template <int NL>
struct SeedAcc {
    using T = HDual<NL>;
    const double* x;
    int nr = 0;
    int start[4], len[4], dir0[4];
    void add_range(int s, int l) {
        if (l <= 0) return;
        dir0[nr] = nr == 0 ? 0 : dir0[nr - 1] + len[nr - 1];
        start[nr] = s; len[nr] = l; ++nr;
    }
    int ndir() const { return nr == 0 ? 0 : dir0[nr - 1] + len[nr - 1]; }
    int global_of(int d) const {
        for (int r = 0; r < nr; ++r)
            if (d >= dir0[r] && d < dir0[r] + len[r]) return start[r] + (d - dir0[r]);
        return -1;
    }
    T operator()(int j) const {
        for (int r = 0; r < nr; ++r)
            if (j >= start[r] && j < start[r] + len[r]) return T::variable(x[j], dir0[r] + (j - start[r]));
        return T(x[j]);
    }
};

// same local seeding as SeedAcc, for the sparsity tracer (pattern_mode 2, ":optimized")
template <int NL>
struct TraceAcc {
    using T = Tracer<NL>;
    const double* x;
    int nr = 0;
    int start[4], len[4], dir0[4];
    void add_range(int s, int l) {
        if (l <= 0) return;
        dir0[nr] = nr == 0 ? 0 : dir0[nr - 1] + len[nr - 1];
        start[nr] = s; len[nr] = l; ++nr;
    }
    int ndir() const { return nr == 0 ? 0 : dir0[nr - 1] + len[nr - 1]; }
    int global_of(int d) const {
        for (int r = 0; r < nr; ++r)
            if (d >= dir0[r] && d < dir0[r] + len[r]) return start[r] + (d - dir0[r]);
        return -1;
    }
    T operator()(int j) const {
        for (int r = 0; r < nr; ++r)
            if (j >= start[r] && j < start[r] + len[r]) return T::variable(0.0, dir0[r] + (j - start[r]));
        return T(0.0);
    }
};

// ---------------------------------------------------------------------------
// The discretised problem.  DOCP struct + scheme constructors:
// src/DOCP_data.jl:176-214 (time), 293-365 (DOCP), trapeze.jl:14-42,
// midpoint.jl:17-39, irk.jl:138-160, irk_stagewise.jl:136-163.
// ---------------------------------------------------------------------------
template <class P, int SCH>
struct Docp {
    using ST = SchemeTraits<SCH>;
    static constexpr int nx = P::nx, nu = P::nu, nv = P::nv, np = P::np, nb = P::nb;
    static constexpr int s = ST::stage;
    static constexpr int control_steps = ST::control_steps;                 // DOCP_data.jl:149
    static constexpr int control_block = ST::stagewise ? nu * s : nu * control_steps;   // irk_stagewise.jl:138, midpoint.jl:20
    static constexpr int step_block = nx + control_block + (ST::irk ? s * nx : 0);
    static constexpr int eqs_block = ST::irk ? nx * (1 + s) : nx;           // irk.jl:144
    static constexpr int cblock = eqs_block + np;
    // local variables touched by one step (x_i, u_i[, k_i], x_{i+1}[, u_{i+1}], v)
    // (implicit Euler: the path rows of node i read the control of block i-1, euler.jl:66-69)
    static constexpr bool prev_control = SCH == EULER_IMPLICIT && np > 0;
    static constexpr int NL_STEP = step_block + nx + (SCH == TRAPEZE ? nu : 0) + nv + (prev_control ? nu : 0);
    static constexpr int NL_NODE = nx + control_block + nv;          // path / trapeze lagrange node
    static constexpr int NL_BC = 2 * nx + nv;                        // boundary + mayer

    int N = 0;                     // time.steps
    std::vector<double> sgrid;     // normalized grid, N+1
    std::vector<double> params;
    ProblemMeta meta;
    Butcher bt;
    int pattern_mode = 0;          // 0 = reference (quirks Q1-Q4 kept), 1 = corrected, 2 = optimized (detected)
    int nvar = 0, ncon = 0;

    Docp(int grid_size, const double* time_grid, int n_time_grid, const double* par, int npar, int pmode) {
        meta = P::meta();
        params = meta.default_params;
        if (par && npar > 0) params.assign(par, par + npar);
        params.resize(P::npar > 0 ? P::npar : 1, 0.0);
        pattern_mode = pmode;
        bt = butcher_for_stage(s);
        // DOCPtime: DOCP_data.jl:176-214
        if (!time_grid) {
            N = grid_size;
            sgrid.resize(N + 1);
            // collect(LinRange(0, 1, N+1)): lerp (1-t)*0 + t*1 with t = j/N
            for (int j = 0; j <= N; ++j) {
                double t = (double)j / (double)N;
                sgrid[j] = (1.0 - t) * 0.0 + t * 1.0;
            }
            if (N == 0) throw std::invalid_argument("grid_size must be >= 1");
        } else {
            for (int j = 0; j + 1 < n_time_grid; ++j)
                if (!(time_grid[j] < time_grid[j + 1]))
                    throw std::invalid_argument("given time grid is not strictly increasing. Aborting...");
            N = n_time_grid - 1;
            if (N < 1) throw std::invalid_argument("time grid needs at least 2 points");
            sgrid.assign(time_grid, time_grid + n_time_grid);
            if (time_grid[0] != 0 || time_grid[N] != 1) {
                double a = time_grid[0], b = time_grid[N];
                for (int j = 0; j <= N; ++j) sgrid[j] = (time_grid[j] - a) / (b - a);
            }
        }
        nvar = N * step_block + nx + nv + (ST::final_control ? nu : 0);
        ncon = N * cblock + np + nb;
    }

    // ---- layout getters: ode/common.jl:113-170, irk_stagewise.jl:173-224 (0-based offsets) ----
    int v_off() const { return nvar - nv; }
    int state_off(int i) const { return (i - 1) * step_block; }
    int control_off(int i, int j = 1) const {  // common.jl:140-155
        if (!ST::final_control && i == N + 1) i = N;
        return (i - 1) * step_block + nx + (j - 1) * nu;
    }
    int stagecontrol_off(int i, int j) const { return (i - 1) * step_block + nx + (j - 1) * nu; }
    int stagevar_off(int i, int j) const { return (i - 1) * step_block + nx + control_block + (j - 1) * nx; }

    template <class Acc> void load(const Acc& a, int off, int n, typename Acc::T* out) const {
        for (int k = 0; k < n; ++k) out[k] = a(off + k);
    }
    template <class Acc> void get_v(const Acc& a, typename Acc::T* v) const { load(a, v_off(), nv, v); }

    // control used at time step i by path constraints / node-based quadrature.
    // generic: common.jl:140-155; stagewise: b-weighted average irk_stagewise.jl:197-205
    template <class Acc> void get_control(const Acc& a, int i, typename Acc::T* u) const {
        using T = typename Acc::T;
        if constexpr (SCH == EULER_IMPLICIT) {
            // euler.jl:66-69: u(]t_i, t_i+1]) = U_i, u(t_1) = U_1  =>  node i reads U_{i-1} (U_1 for i = 1)
            if (i == 1) i = 2;
            load(a, (i - 2) * step_block + nx, nu, u);
        } else if constexpr (ST::stagewise) {
            if (i == N + 1) i = N;
            for (int k = 0; k < nu; ++k) u[k] = bt.b[0] * a(stagecontrol_off(i, 1) + k);
            for (int j = 2; j <= s; ++j)
                for (int k = 0; k < nu; ++k) u[k] = u[k] + bt.b[j - 1] * a(stagecontrol_off(i, j) + k);
            (void)sizeof(T);
        } else {
            load(a, control_off(i), nu, u);
        }
    }

    // get_time_grid: DOCP_data.jl:437-458  (t_i = t0 + s_i * (tf - t0)); also the fixed grid l.210
    template <class Acc> typename Acc::T time_at(const Acc& a, int i) const {
        using T = typename Acc::T;
        T v[nv > 0 ? nv : 1];
        get_v(a, v);
        T t0 = P::template t0<T>(v, params.data());
        T tf = P::template tf<T>(v, params.data());
        return t0 + sgrid[i - 1] * (tf - t0);
    }

    // ---------------- state / stage equations of step i (rows 1..eqs_block of the step) -------------
    template <class Acc> void step_state_eqs(const Acc& a, int i, typename Acc::T* c) const {
        using T = typename Acc::T;
        const double* p = params.data();
        T v[nv > 0 ? nv : 1];
        get_v(a, v);
        T ti = time_at(a, i), tip1 = time_at(a, i + 1);
        T xi[nx], xip1[nx];
        load(a, state_off(i), nx, xi);
        load(a, state_off(i + 1), nx, xip1);
        if constexpr (ST::midpoint) {
            // setWorkArray midpoint.jl:47-72 + stepStateConstraints! midpoint.jl:124-156
            T ts = 0.5 * (ti + tip1);
            T xs[nx], u[nu > 0 ? nu : 1], f[nx];
            for (int k = 0; k < nx; ++k) xs[k] = 0.5 * (xi[k] + xip1[k]);
            T hi = (tip1 - ti) / (double)control_steps;
            if constexpr (control_steps == 1) {
                load(a, control_off(i), nu, u);
                P::template dynamics<T>(f, ts, xs, u, v, p);
                for (int k = 0; k < nx; ++k) c[k] = xip1[k] - (xi[k] + hi * f[k]);
            } else {
                // midpoint.jl:145-154: every sub-step uses the dynamics at (ts, xs) of the WHOLE step (the work array
                // of l.57-65) with its own control; x_next = xi; x_next += hi * work_j
                T x_next[nx];
                for (int k = 0; k < nx; ++k) x_next[k] = xi[k];
                for (int j = 1; j <= control_steps; ++j) {
                    load(a, control_off(i, j), nu, u);
                    P::template dynamics<T>(f, ts, xs, u, v, p);
                    for (int k = 0; k < nx; ++k) x_next[k] = x_next[k] + hi * f[k];
                }
                for (int k = 0; k < nx; ++k) c[k] = xip1[k] - x_next[k];
            }
        } else if constexpr (SCH == EULER_EXPLICIT || SCH == EULER_IMPLICIT) {
            // setWorkArray euler.jl:79-104 (f at node i, or at node i+1 for the implicit rule) +
            // stepStateConstraints! euler.jl:140-159
            T u[nu > 0 ? nu : 1], f[nx];
            if constexpr (SCH == EULER_EXPLICIT) {
                get_control(a, i, u);
                P::template dynamics<T>(f, ti, xi, u, v, p);
            } else {
                get_control(a, i + 1, u);
                P::template dynamics<T>(f, tip1, xip1, u, v, p);
            }
            T hi = tip1 - ti;
            for (int k = 0; k < nx; ++k) c[k] = xip1[k] - (xi[k] + hi * f[k]);
        } else if constexpr (SCH == TRAPEZE) {
            // setWorkArray trapeze.jl:50-71 + stepStateConstraints! trapeze.jl:118-142
            T ui[nu > 0 ? nu : 1], uip1[nu > 0 ? nu : 1], fi[nx], fip1[nx];
            load(a, control_off(i), nu, ui);
            load(a, control_off(i + 1), nu, uip1);
            P::template dynamics<T>(fi, ti, xi, ui, v, p);
            P::template dynamics<T>(fip1, tip1, xip1, uip1, v, p);
            T half_hi = 0.5 * (tip1 - ti);
            for (int k = 0; k < nx; ++k) {
                T x_next = xi[k];
                x_next = x_next + half_hi * (fi[k] + fip1[k]);
                c[k] = xip1[k] - x_next;
            }
        } else {
            // generic IRK: irk.jl:236-308 ; stagewise: irk_stagewise.jl:394-460
            T hi = tip1 - ti;
            T sumbk[nx], xij[nx], kij[nx], u[nu > 0 ? nu : 1], f[nx];
            if constexpr (!ST::stagewise) load(a, control_off(i), nu, u);
            int offset_stage = nx;
            for (int j = 1; j <= s; ++j) {
                T tij = ti + bt.c[j - 1] * hi;
                load(a, stagevar_off(i, j), nx, kij);
                if constexpr (ST::stagewise) load(a, stagecontrol_off(i, j), nu, u);
                for (int k = 0; k < nx; ++k) {
                    if (j == 1) sumbk[k] = bt.b[0] * kij[k];
                    else sumbk[k] = sumbk[k] + bt.b[j - 1] * kij[k];
                }
                for (int k = 0; k < nx; ++k) xij[k] = xi[k];
                for (int l = 1; l <= s; ++l) {
                    T kil[nx];
                    load(a, stagevar_off(i, l), nx, kil);
                    for (int k = 0; k < nx; ++k) xij[k] = xij[k] + (hi * bt.a[j - 1][l - 1]) * kil[k];
                }
                P::template dynamics<T>(f, tij, xij, u, v, p);
                for (int k = 0; k < nx; ++k) c[offset_stage + k] = kij[k] - f[k];
                offset_stage += nx;
            }
            for (int k = 0; k < nx; ++k) c[k] = xip1[k] - (xi[k] + hi * sumbk[k]);
        }
    }

    // stepPathConstraints!: DOCP_functions.jl:122-140 (node i in 1..N+1)
    template <class Acc> void path_at(const Acc& a, int i, typename Acc::T* c) const {
        using T = typename Acc::T;
        if constexpr (np > 0) {
            T v[nv > 0 ? nv : 1], xi[nx], ui[nu > 0 ? nu : 1];
            get_v(a, v);
            T ti = time_at(a, i);
            load(a, state_off(i), nx, xi);
            get_control(a, i, ui);
            P::template path<T>(c, ti, xi, ui, v, params.data());
        } else {
            (void)a; (void)i; (void)c;
        }
    }

    template <class Acc> void boundary(const Acc& a, typename Acc::T* c) const {
        using T = typename Acc::T;
        if constexpr (nb > 0) {
            T v[nv > 0 ? nv : 1], x0[nx], xf[nx];
            get_v(a, v);
            load(a, state_off(1), nx, x0);
            load(a, state_off(N + 1), nx, xf);
            P::template boundary<T>(c, x0, xf, v, params.data());
        }
    }

    // __constraints!: DOCP_functions.jl:80-115
    void constraints(const double* xu, double* c) const {
        PlainAcc a{xu};
        for (int i = 1; i <= N; ++i) {
            int off = (i - 1) * cblock;
            step_state_eqs(a, i, c + off);
            if (np > 0) path_at(a, i, c + off + eqs_block);
        }
        if (np > 0) path_at(a, N + 1, c + N * cblock);
        if (nb > 0) boundary(a, c + ncon - nb);
    }

    // ----- running cost: common.jl:242-244 -> integral() per scheme ----------------------------
    // lagrange term attached to step i (midpoint.jl:87-97, irk.jl:189-225, irk_stagewise.jl:351-381)
    template <class Acc> typename Acc::T lagrange_step(const Acc& a, int i) const {
        using T = typename Acc::T;
        const double* p = params.data();
        T v[nv > 0 ? nv : 1];
        get_v(a, v);
        T ti = time_at(a, i), tip1 = time_at(a, i + 1);
        T xi[nx];
        load(a, state_off(i), nx, xi);
        if constexpr (ST::midpoint) {
            T xip1[nx], xs[nx], u[nu > 0 ? nu : 1];
            load(a, state_off(i + 1), nx, xip1);
            for (int k = 0; k < nx; ++k) xs[k] = 0.5 * (xi[k] + xip1[k]);
            if constexpr (control_steps == 1) {
                T hi = tip1 - ti;
                T ts = 0.5 * (ti + tip1);
                load(a, control_off(i), nu, u);
                return hi * P::template lagrange<T>(ts, xs, u, v, p);
            } else {
                // integral(): midpoint.jl:98-113 -- sub-step midpoints tij = ti + (j - 0.5) hi
                T hi = (tip1 - ti) / (double)control_steps;
                T value(0.0);
                for (int j = 1; j <= control_steps; ++j) {
                    T tij = ti + (j - 0.5) * hi;
                    load(a, control_off(i, j), nu, u);
                    T term = hi * P::template lagrange<T>(tij, xs, u, v, p);
                    if (j == 1) value = term;
                    else value = value + term;
                }
                return value;
            }
        } else if constexpr (SCH == EULER_EXPLICIT || SCH == EULER_IMPLICIT) {
            // integral(): euler.jl:111-133
            T hi = tip1 - ti;
            T u[nu > 0 ? nu : 1];
            if constexpr (SCH == EULER_EXPLICIT) {
                get_control(a, i, u);
                return hi * P::template lagrange<T>(ti, xi, u, v, p);
            } else {
                T xip1[nx];
                load(a, state_off(i + 1), nx, xip1);
                get_control(a, i + 1, u);
                return hi * P::template lagrange<T>(tip1, xip1, u, v, p);
            }
        } else {
            T hi = tip1 - ti;
            T u[nu > 0 ? nu : 1], xij[nx];
            if constexpr (!ST::stagewise) load(a, control_off(i), nu, u);
            T local_sum(0.0);
            for (int j = 1; j <= s; ++j) {
                T tij = ti + bt.c[j - 1] * hi;
                if constexpr (ST::stagewise) load(a, stagecontrol_off(i, j), nu, u);
                for (int k = 0; k < nx; ++k) xij[k] = xi[k];
                for (int l = 1; l <= s; ++l) {
                    T kil[nx];
                    load(a, stagevar_off(i, l), nx, kil);
                    for (int k = 0; k < nx; ++k) xij[k] = xij[k] + (hi * bt.a[j - 1][l - 1]) * kil[k];
                }
                T term = bt.b[j - 1] * P::template lagrange<T>(tij, xij, u, v, p);
                if (j == 1) local_sum = term;
                else local_sum = local_sum + term;
            }
            return hi * local_sum;
        }
    }
    // trapeze node term: trapeze.jl:78-110 (weights h_1/2, (t_{i+1}-t_{i-1})/2, h_N/2)
    template <class Acc> typename Acc::T lagrange_node_trapeze(const Acc& a, int i) const {
        using T = typename Acc::T;
        T v[nv > 0 ? nv : 1], xi[nx], ui[nu > 0 ? nu : 1];
        get_v(a, v);
        T ti = time_at(a, i);
        load(a, state_off(i), nx, xi);
        load(a, control_off(i), nu, ui);
        T w;
        if (i == 1) w = time_at(a, 2) - time_at(a, 1);
        else if (i == N + 1) w = time_at(a, N + 1) - time_at(a, N);
        else w = time_at(a, i + 1) - time_at(a, i - 1);
        return w / 2.0 * P::template lagrange<T>(ti, xi, ui, v, params.data());
    }
    template <class Acc> typename Acc::T mayer(const Acc& a) const {
        using T = typename Acc::T;
        T v[nv > 0 ? nv : 1], x0[nx], xf[nx];
        get_v(a, v);
        load(a, state_off(1), nx, x0);
        load(a, state_off(N + 1), nx, xf);
        return P::template mayer<T>(x0, xf, v, params.data());
    }

    // __objective: DOCP_functions.jl:23-54 (sequential left-to-right accumulation)
    double objective(const double* xu) const {
        PlainAcc a{xu};
        double obj_mayer = P::has_mayer ? mayer(a) : 0.0;
        double obj_lagrange = 0.0;
        if (P::has_lagrange) {
            double value = 0.0;
            if constexpr (SCH == TRAPEZE) {
                for (int i = 1; i <= N + 1; ++i) value += lagrange_node_trapeze(a, i);
            } else {
                for (int i = 1; i <= N; ++i) value += lagrange_step(a, i);
            }
            obj_lagrange = value;
        }
        return obj_mayer + obj_lagrange;
    }

    // ----- bounds: DOCP_functions.jl:163-191, DOCP_variables.jl:21-63,88-98, irk_stagewise.jl:250-300 -----
    void constraint_bounds(double* lb, double* ub) const {
        for (int k = 0; k < ncon; ++k) lb[k] = ub[k] = 0.0;
        int offset = 0;
        for (int i = 1; i <= N + 1; ++i) {
            if (i <= N) offset += eqs_block;
            for (int k = 0; k < np; ++k) {
                lb[offset + k] = meta.path_lb[k];
                ub[offset + k] = meta.path_ub[k];
            }
            offset += np;
        }
        for (int k = 0; k < nb; ++k) {
            lb[offset + k] = meta.bc_lb[k];
            ub[offset + k] = meta.bc_ub[k];
        }
    }
    static void bounds_block(int dim, const std::vector<Box>& box, std::vector<double>& lb, std::vector<double>& ub) {
        lb.assign(dim, -INF);
        ub.assign(dim, INF);
        for (auto& b : box) { lb[b.idx - 1] = b.lb; ub[b.idx - 1] = b.ub; }
    }
    void variable_bounds(double* lb, double* ub) const {
        for (int k = 0; k < nvar; ++k) { lb[k] = -INF; ub[k] = INF; }
        std::vector<double> xl, xub, ul, uub, vl, vub;
        bounds_block(nx, meta.state_box, xl, xub);
        bounds_block(nu, meta.control_box, ul, uub);
        for (int i = 1; i <= N + 1; ++i) {
            for (int k = 0; k < nx; ++k) { lb[state_off(i) + k] = xl[k]; ub[state_off(i) + k] = xub[k]; }
            if (ST::stagewise) continue;
            // set_control_at_time_step!: common.jl:212-225 (final control only if the scheme has one)
            if (nu > 0 && (i <= N || ST::final_control))
                for (int j = 1; j <= control_steps; ++j)      // DOCP_variables.jl:44
                    for (int k = 0; k < nu; ++k) { lb[control_off(i, j) + k] = ul[k]; ub[control_off(i, j) + k] = uub[k]; }
        }
        if (ST::stagewise && nu > 0)
            for (int i = 1; i <= N; ++i)
                for (int j = 1; j <= s; ++j)
                    for (int k = 0; k < nu; ++k) { lb[stagecontrol_off(i, j) + k] = ul[k]; ub[stagecontrol_off(i, j) + k] = uub[k]; }
        if (nv > 0) {
            bounds_block(nv, meta.variable_box, vl, vub);
            for (int k = 0; k < nv; ++k) { lb[v_off() + k] = vl[k]; ub[v_off() + k] = vub[k]; }
        }
    }
    // __initial_guess: DOCP_variables.jl:122-145 (0.1 everywhere, then constant user init at nodes);
    // stagewise: irk_stagewise.jl:302-335 (controls at every stage)
    void initial_guess(double* x0) const {
        for (int k = 0; k < nvar; ++k) x0[k] = 0.1;
        if (!meta.init_variable.empty())
            for (int k = 0; k < nv; ++k) x0[v_off() + k] = meta.init_variable[k];
        for (int i = 1; i <= N + 1; ++i) {
            if (!meta.init_state.empty())
                for (int k = 0; k < nx; ++k) x0[state_off(i) + k] = meta.init_state[k];
            if (!ST::stagewise && nu > 0 && !meta.init_control.empty() && (i <= N || ST::final_control))
                for (int j = 1; j <= control_steps; ++j)      // DOCP_variables.jl:138
                    for (int k = 0; k < nu; ++k) x0[control_off(i, j) + k] = meta.init_control[k];
        }
        if (ST::stagewise && nu > 0 && !meta.init_control.empty())
            for (int i = 1; i <= N; ++i)
                for (int j = 1; j <= s; ++j)
                    for (int k = 0; k < nu; ++k) x0[stagecontrol_off(i, j) + k] = meta.init_control[k];
    }

    // ----- sparsity patterns (1-based, as the reference builds them) ---------------------------
    // trapeze.jl:149-233, midpoint.jl:163-233, irk.jl:315-416, irk_stagewise.jl:468-558
    // ----- ":optimized" patterns: structural detection by tracing the restated functions piece by piece,
    // the way ADNLPModels' default backend traces the reference's closures (collocation.jl:131-134) -----
    template <int NL>
    void trace_rows(const TraceAcc<NL>& a, const Tracer<NL>* c, int nrows, int64_t row0, TripletList* J) const {
        if (!J) return;
        for (int r = 0; r < nrows; ++r)
            for (int d = 0; d < a.ndir(); ++d)
                if (c[r].dep(d)) J->add(row0 + r + 1, a.global_of(d) + 1);
    }
    template <int NL>
    void trace_lag(const TraceAcc<NL>& a, const Tracer<NL>& l, TripletList* H) const {
        if (!H) return;
        for (int d1 = 0; d1 < a.ndir(); ++d1)
            for (int d2 = 0; d2 < a.ndir(); ++d2)
                if (l.pair(d1, d2)) H->add(a.global_of(d1) + 1, a.global_of(d2) + 1);
    }
    void detect(TripletList* J, TripletList* H) const {
        for (int i = 1; i <= N; ++i) {
            TraceAcc<NL_STEP> a;
            a.x = nullptr;
            if (prev_control && i >= 2) a.add_range(control_off(i - 1), nu);
            a.add_range(state_off(i), step_block + nx + (SCH == TRAPEZE ? nu : 0));
            a.add_range(v_off(), nv);
            Tracer<NL_STEP> c[cblock > 0 ? cblock : 1];
            step_state_eqs(a, i, c);
            if (np > 0) path_at(a, i, c + eqs_block);
            trace_rows(a, c, cblock, (int64_t)(i - 1) * cblock, J);
            Tracer<NL_STEP> l(0.0);
            for (int r = 0; r < cblock; ++r) l = l + c[r];
            if (P::has_lagrange && SCH != TRAPEZE) l = l + lagrange_step(a, i);
            trace_lag(a, l, H);
        }
        if constexpr (SCH == TRAPEZE) {
            if (P::has_lagrange)
                for (int i = 1; i <= N + 1; ++i) {
                    TraceAcc<NL_NODE> a;
                    a.x = nullptr;
                    a.add_range(state_off(i), nx + nu);
                    a.add_range(v_off(), nv);
                    trace_lag(a, lagrange_node_trapeze(a, i), H);
                }
        }
        if constexpr (np > 0) {
            TraceAcc<NL_NODE> a;
            a.x = nullptr;
            if (SCH == TRAPEZE) a.add_range(state_off(N + 1), nx + nu);
            else {
                a.add_range(control_off(N), control_block);
                a.add_range(state_off(N + 1), nx);
            }
            a.add_range(v_off(), nv);
            Tracer<NL_NODE> c[np];
            path_at(a, N + 1, c);
            trace_rows(a, c, np, (int64_t)N * cblock, J);
            Tracer<NL_NODE> l(0.0);
            for (int r = 0; r < np; ++r) l = l + c[r];
            trace_lag(a, l, H);
        }
        {
            TraceAcc<NL_BC> a;
            a.x = nullptr;
            a.add_range(state_off(1), nx);
            a.add_range(state_off(N + 1), nx);
            a.add_range(v_off(), nv);
            Tracer<NL_BC> l(0.0);
            if constexpr (nb > 0) {
                Tracer<NL_BC> c[nb];
                boundary(a, c);
                trace_rows(a, c, nb, ncon - nb, J);
                for (int r = 0; r < nb; ++r) l = l + c[r];
            }
            if (P::has_mayer) l = l + mayer(a);
            trace_lag(a, l, H);
        }
    }

    Pattern jacobian_pattern() const {
        TripletList L;
        if (pattern_mode == 2) {
            detect(&L, nullptr);
            return L.build(ncon, nvar, false);
        }
        const int64_t v_start = nvar - nv + 1, v_end = nvar;
        const int64_t B = step_block;
        for (int64_t i = 1; i <= N; ++i) {
            const int64_t c_offset = (i - 1) * cblock;
            const int64_t var_offset = (i - 1) * B;
            const int64_t dyn_start = c_offset + 1, dyn_end = c_offset + nx;
            if constexpr (ST::midpoint) {
                const int64_t xi_start = var_offset + 1, ui_end = var_offset + nx + nu;
                const int64_t xip1_end = var_offset + B + nx;
                L.add_block(dyn_start, dyn_end, xi_start, xip1_end);
                L.add_block(c_offset + nx + 1, c_offset + cblock, xi_start, ui_end);
                L.add_block(c_offset + 1, c_offset + cblock, v_start, v_end);
            } else if constexpr (SCH == EULER_EXPLICIT || SCH == EULER_IMPLICIT) {
                // euler.jl:166-262 (the Lagrange-cost lines only repeat entries of the dynamics block)
                const int64_t path_start = c_offset + nx + 1, path_end = c_offset + cblock;
                const int64_t xi_start = var_offset + 1, xi_end = var_offset + nx;
                const int64_t ui_start = var_offset + nx + 1, ui_end = var_offset + nx + nu;
                const int64_t xip1_end = var_offset + B + nx;
                L.add_block(dyn_start, dyn_end, xi_start, xi_end);
                L.add_block(dyn_start, dyn_end, ui_start, xip1_end);
                L.add_block(dyn_start, dyn_end, v_start, v_end);
                L.add_block(path_start, path_end, xi_start, xi_end);
                if (SCH == EULER_IMPLICIT && pattern_mode == 1 && i >= 2)
                    // node i of the implicit rule reads the control of block i-1 (euler.jl:66-69); the reference's
                    // :manual pattern lists block i (euler.jl:214-216)
                    L.add_block(path_start, path_end, ui_start - B, ui_end - B);
                else
                    L.add_block(path_start, path_end, ui_start, ui_end);
                L.add_block(path_start, path_end, v_start, v_end);
            } else if constexpr (SCH == TRAPEZE) {
                const int64_t path_start = c_offset + nx + 1, path_end = c_offset + cblock;
                const int64_t xi_start = var_offset + 1, xi_end = var_offset + nx;
                const int64_t ui_start = var_offset + nx + 1, ui_end = var_offset + nx + nu;
                const int64_t xip1_end = var_offset + nx + nu + nx;
                const int64_t uip1_start = var_offset + nx * 2 + nu + 1, uip1_end = var_offset + nx * 2 + nu * 2;
                L.add_block(dyn_start, dyn_end, xi_start, xi_end);
                L.add_block(dyn_start, dyn_end, ui_start, xip1_end);
                L.add_block(dyn_start, dyn_end, uip1_start, uip1_end);
                L.add_block(path_start, path_end, xi_start, xi_end);
                L.add_block(path_start, path_end, ui_start, ui_end);
                // trapeze.jl:203 -- only the PATH rows get the v column (quirk Q1)
                L.add_block(path_start, path_end, v_start, v_end);
                if (pattern_mode == 1) L.add_block(dyn_start, dyn_end, v_start, v_end);
            } else {
                const int64_t stage_start = c_offset + nx + 1, stage_end = c_offset + (s + 1) * nx;
                const int64_t path_start = c_offset + (s + 1) * nx + 1, path_end = c_offset + cblock;
                const int64_t xi_start = var_offset + 1, xi_end = var_offset + nx;
                const int64_t ui_end = var_offset + nx + control_block;
                const int64_t ki_start = var_offset + nx + control_block + 1, ki_end = var_offset + B;
                const int64_t xip1_end = var_offset + B + nx;
                L.add_block(dyn_start, dyn_end, xi_start, xi_end);
                L.add_block(dyn_start, dyn_end, ki_start, xip1_end);
                L.add_block(dyn_start, dyn_end, v_start, v_end);
                L.add_block(stage_start, stage_end, xi_start, ki_end);
                L.add_block(stage_start, stage_end, v_start, v_end);
                L.add_block(path_start, path_end, xi_start, ui_end);
                L.add_block(path_start, path_end, v_start, v_end);
            }
        }
        // final path constraints (xf, uf, v)
        {
            const int64_t c_offset = (int64_t)N * cblock;
            const int64_t var_offset = (int64_t)N * B;
            const int64_t xf_start = var_offset + 1, xf_end = var_offset + nx;
            int64_t uf_start, uf_end;
            if (SCH == TRAPEZE) { uf_start = var_offset + nx + 1; uf_end = var_offset + nx + nu; }
            else { uf_start = var_offset - B + nx + 1; uf_end = var_offset - B + nx + control_block; }
            L.add_block(c_offset + 1, c_offset + np, xf_start, xf_end);
            L.add_block(c_offset + 1, c_offset + np, uf_start, uf_end);
            L.add_block(c_offset + 1, c_offset + np, v_start, v_end);
            // boundary constraints (x0, xf, v)
            const int64_t b_offset = c_offset + np;
            L.add_block(b_offset + 1, b_offset + nb, 1, nx);
            L.add_block(b_offset + 1, b_offset + nb, xf_start, xf_end);
            L.add_block(b_offset + 1, b_offset + nb, v_start, v_end);
        }
        // irk_stagewise.jl:550-552, euler.jl:257-259: stray entry (quirk Q3)
        if ((ST::stagewise || SCH == EULER_EXPLICIT || SCH == EULER_IMPLICIT) && P::has_lagrange && pattern_mode == 0)
            L.add(ncon, nx);
        return L.build(ncon, nvar, false);
    }

    // trapeze.jl:240-303, midpoint.jl:240-300, irk.jl:423-496, irk_stagewise.jl:565-638; then tril [3P]
    Pattern hessian_pattern() const {
        TripletList L;
        if (pattern_mode == 2) {
            detect(nullptr, &L);
            return L.build(nvar, nvar, true);
        }
        const int64_t v_start = nvar - nv + 1, v_end = nvar;
        const int64_t B = step_block;
        L.add_block(v_start, v_end, v_start, v_end);
        for (int64_t i = 1; i <= N; ++i) {
            const int64_t var_offset = (i - 1) * B;
            if constexpr (ST::midpoint) {
                const int64_t xi_start = var_offset + 1, xip1_end = var_offset + B + nx;
                L.add_block(xi_start, xip1_end, xi_start, xip1_end);
                L.add_block(xi_start, xip1_end, v_start, v_end, true);
            } else if constexpr (SCH == EULER_EXPLICIT || SCH == EULER_IMPLICIT) {
                // euler.jl:269-355
                const int64_t xi_start = var_offset + 1, xi_end = var_offset + nx;
                const int64_t xip1_start = var_offset + B + 1, xip1_end = var_offset + B + nx;
                const int64_t ui_start = var_offset + nx + 1, ui_end = var_offset + nx + nu;
                L.add_block(ui_start, ui_end, ui_start, ui_end);
                L.add_block(ui_start, ui_end, v_start, v_end, true);
                if (SCH == EULER_IMPLICIT) {
                    L.add_block(xip1_start, xip1_end, xip1_start, xip1_end);
                    L.add_block(xip1_start, xip1_end, ui_start, ui_end, true);
                    L.add_block(xip1_start, xip1_end, v_start, v_end, true);
                }
                L.add_block(xi_start, xi_end, xi_start, xi_end);
                L.add_block(xi_start, xi_end, ui_start, ui_end, true);
                L.add_block(xi_start, xi_end, v_start, v_end, true);
            } else if constexpr (SCH == TRAPEZE) {
                const int64_t var_block = B * 2;
                L.add_block(var_offset + 1, var_offset + var_block, var_offset + 1, var_offset + var_block);
                L.add_block(var_offset + 1, var_offset + var_block, v_start, v_end, true);
            } else {
                const int64_t xi_start = var_offset + 1, ki_end = var_offset + B;
                L.add_block(xi_start, ki_end, xi_start, ki_end);
                L.add_block(xi_start, ki_end, v_start, v_end, true);
            }
        }
        const int64_t var_offset = (int64_t)N * B;
        const int64_t xf_start = var_offset + 1, xf_end = var_offset + nx;
        if constexpr (ST::irk) {
            const int64_t uf_start = var_offset - B + nx + 1, uf_end = var_offset - B + nx + control_block;
            L.add_block(xf_start, xf_end, xf_start, xf_end);
            L.add_block(uf_start, uf_end, uf_start, uf_end);
            L.add_block(xf_start, xf_end, uf_start, uf_end, true);
            if constexpr (ST::stagewise) {
                // irk_stagewise.jl:624: range xf_start:uf_end is EMPTY (quirk Q2)
                L.add_block(xf_start, uf_end, v_start, v_end, true);
                if (pattern_mode == 1) L.add_block(xf_start, xf_end, v_start, v_end, true);
            } else {
                L.add_block(xf_start, xf_end, v_start, v_end, true);
            }
            L.add_block(uf_start, uf_end, v_start, v_end, true);
            L.add_block(1, nx, xf_start, xf_end, true);
        } else if constexpr (ST::midpoint || SCH == EULER_EXPLICIT || SCH == EULER_IMPLICIT) {
            L.add_block(1, nx, xf_start, xf_end, true);  // always (quirk Q4): midpoint.jl:286-291, euler.jl:345-350
        } else {
            if (P::has_mayer || nb > 0) L.add_block(1, nx, xf_start, xf_end, true);
        }
        return L.build(nvar, nvar, true);
    }

    // ----- derivative values by element-wise second-order forward AD -------------------------
    // The Lagrangian l(x) = sigma*f(x) + y'c(x) (NLPModels convention [3P]) is a sum of pieces that
    // each touch few variables: one per step, one per path node, boundary+mayer.  Each piece is
    // evaluated with hyper-duals seeded on its own variables through the restated functions above.
    struct DerivOut {
        const Pattern* jp = nullptr; double* jvals = nullptr;   // may be null
        const Pattern* hp = nullptr; double* hvals = nullptr;   // may be null
        double* grad = nullptr;                                 // may be null (dense, nvar)
        int64_t dropped_j = 0, dropped_h = 0;                    // nonzeros outside the pattern
    };

    template <int NL>
    void scatter_rows(const SeedAcc<NL>& a, const HDual<NL>* c, int nrows, int64_t row0, DerivOut& o) const {
        if (!o.jvals) return;
        const int nd = a.ndir();
        for (int r = 0; r < nrows; ++r)
            for (int d = 0; d < nd; ++d) {
                int64_t pos = o.jp->find(row0 + r + 1, a.global_of(d) + 1);
                if (pos >= 0) o.jvals[pos] = c[r].g[d];
                else if (c[r].g[d] != 0.0) {
#pragma omp atomic
                    o.dropped_j++;
                }
            }
    }
    template <int NL>
    void scatter_lag(const SeedAcc<NL>& a, const HDual<NL>& l, DerivOut& o) const {
        if (!o.hvals) return;
        const int nd = a.ndir();
        for (int d1 = 0; d1 < nd; ++d1)
            for (int d2 = 0; d2 <= d1; ++d2) {
                int64_t g1 = a.global_of(d1) + 1, g2 = a.global_of(d2) + 1;
                int64_t pos = o.hp->find(std::max(g1, g2), std::min(g1, g2));
                double val = l.hess(d1, d2);
                if (pos >= 0) {
                    if (std::min(g1, g2) > nvar - nv) {
                        // (v,v) entries are shared by every piece
#pragma omp atomic
                        o.hvals[pos] += val;
                    } else {
                        o.hvals[pos] += val;
                    }
                } else if (val != 0.0) {
#pragma omp atomic
                    o.dropped_h++;
                }
            }
    }
    template <int NL>
    void scatter_grad(const SeedAcc<NL>& a, const HDual<NL>& f, DerivOut& o) const {
        if (!o.grad) return;
        for (int d = 0; d < a.ndir(); ++d) {
            int g = a.global_of(d);
            if (g >= nvar - nv) {
#pragma omp atomic
                o.grad[g] += f.g[d];
            } else {
                o.grad[g] += f.g[d];
            }
        }
    }

    void step_piece(const double* xu, const double* y, double sigma, int i, DerivOut& o) const {
        using T = HDual<NL_STEP>;
        SeedAcc<NL_STEP> a;
        a.x = xu;
        if (prev_control && i >= 2) a.add_range(control_off(i - 1), nu);
        a.add_range(state_off(i), step_block + nx + (SCH == TRAPEZE ? nu : 0));
        a.add_range(v_off(), nv);
        T c[cblock > 0 ? cblock : 1];
        step_state_eqs(a, i, c);
        if (np > 0) path_at(a, i, c + eqs_block);
        const int64_t row0 = (int64_t)(i - 1) * cblock;
        scatter_rows(a, c, cblock, row0, o);
        T l(0.0);
        if (y) for (int r = 0; r < cblock; ++r) l = l + y[row0 + r] * c[r];
        if (P::has_lagrange && SCH != TRAPEZE) {
            T lc = lagrange_step(a, i);
            scatter_grad(a, lc, o);
            l = l + sigma * lc;
        }
        scatter_lag(a, l, o);
    }
    void node_piece_trapeze_lagrange(const double* xu, double sigma, int i, DerivOut& o) const {
        if constexpr (SCH == TRAPEZE) {
            using T = HDual<NL_NODE>;
            SeedAcc<NL_NODE> a;
            a.x = xu;
            a.add_range(state_off(i), nx + nu);
            a.add_range(v_off(), nv);
            T lc = lagrange_node_trapeze(a, i);
            scatter_grad(a, lc, o);
            scatter_lag(a, sigma * lc, o);
        }
    }
    void final_path_piece(const double* xu, const double* y, DerivOut& o) const {
        if constexpr (np > 0) {
            using T = HDual<NL_NODE>;
            SeedAcc<NL_NODE> a;
            a.x = xu;
            if (SCH == TRAPEZE) a.add_range(state_off(N + 1), nx + nu);
            else {
                a.add_range(control_off(N), control_block);
                a.add_range(state_off(N + 1), nx);
            }
            a.add_range(v_off(), nv);
            T c[np];
            path_at(a, N + 1, c);
            const int64_t row0 = (int64_t)N * cblock;
            scatter_rows(a, c, np, row0, o);
            T l(0.0);
            if (y) for (int r = 0; r < np; ++r) l = l + y[row0 + r] * c[r];
            scatter_lag(a, l, o);
        }
    }
    void boundary_piece(const double* xu, const double* y, double sigma, DerivOut& o) const {
        using T = HDual<NL_BC>;
        SeedAcc<NL_BC> a;
        a.x = xu;
        a.add_range(state_off(1), nx);
        a.add_range(state_off(N + 1), nx);
        a.add_range(v_off(), nv);
        T l(0.0);
        if constexpr (nb > 0) {
            T c[nb];
            boundary(a, c);
            const int64_t row0 = ncon - nb;
            scatter_rows(a, c, nb, row0, o);
            if (y) for (int r = 0; r < nb; ++r) l = l + y[row0 + r] * c[r];
        }
        if (P::has_mayer) {
            T m = mayer(a);
            scatter_grad(a, m, o);
            l = l + sigma * m;
        }
        scatter_lag(a, l, o);
    }

    void derivatives(const double* xu, const double* y, double sigma, DerivOut& o) const {
        if (o.jvals) std::fill(o.jvals, o.jvals + o.jp->nnz(), 0.0);
        if (o.hvals) std::fill(o.hvals, o.hvals + o.hp->nnz(), 0.0);
        if (o.grad) std::fill(o.grad, o.grad + nvar, 0.0);
        // steps of equal parity touch disjoint variable blocks, so each parity class may run in
        // parallel (OpenMP, used only by the timed CPU baseline); results do not depend on threads
        // except for the summation order of the (v,v) entries.
        for (int parity = 1; parity >= 0; --parity) {
#pragma omp parallel for schedule(static)
            for (int i = 2 - parity; i <= N; i += 2) step_piece(xu, y, sigma, i, o);
        }
        if (P::has_lagrange && SCH == TRAPEZE) {
#pragma omp parallel for schedule(static)
            for (int i = 1; i <= N + 1; ++i) node_piece_trapeze_lagrange(xu, sigma, i, o);
        }
        final_path_piece(xu, y, o);
        boundary_piece(xu, y, sigma, o);
    }
};

}  // namespace oracle
