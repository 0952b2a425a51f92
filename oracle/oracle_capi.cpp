// TEST INFRASTRUCTURE ONLY -- C entry points of the CPU oracle (loaded with ctypes from
// oracle/oracle.py by tests/, __graft_entry__.smoke() and bench.py's CPU legs).
// The product (libctdirect_b200.so) never links or loads this file.
#include <cstdio>
#include <memory>
#include <string>

#include "docp_oracle.hpp"

#ifdef _OPENMP
#include <omp.h>
#endif

using namespace oracle;

namespace {

struct ModelBase {
    virtual ~ModelBase() {}
    virtual void dims(int64_t* out) = 0;  // nvar, ncon, nnzj, nnzh, nx, nu, nv, np, nb, N, step_block, cblock
    virtual void flags(int* out) = 0;     // free_t0, free_tf, lagrange, mayer, max
    virtual void structure(int which, int64_t* rows, int64_t* cols) = 0;
    virtual void bounds(double* lvar, double* uvar, double* lcon, double* ucon) = 0;
    virtual void x0(double* x) = 0;
    virtual void time_grid(const double* x, double* t) = 0;
    virtual double obj(const double* x) = 0;
    virtual void cons(const double* x, double* c) = 0;
    virtual void derivs(const double* x, const double* y, double sigma, double* grad, double* jvals,
                        double* hvals, int64_t* dropped) = 0;
    virtual int control_steps() const = 0;
    virtual void control_time(const double* x, double* tc) = 0;
    virtual void unpack(const double* x, const double* y, double* time, double* state, double* control,
                        double* variable, double* costate, double* path_dual, double* boundary_dual) = 0;
    virtual void pack(const double* state, const double* control, const double* variable,
                      const double* stage_control, double* x) = 0;
};

template <class P, int SCH>
struct Model : ModelBase {
    Docp<P, SCH> d;
    Pattern jp, hp;
    Model(int N, const double* tg, int ntg, const double* par, int npar, int pmode)
        : d(N, tg, ntg, par, npar, pmode) {
        jp = d.jacobian_pattern();
        hp = d.hessian_pattern();
    }
    void dims(int64_t* o) override {
        o[0] = d.nvar; o[1] = d.ncon; o[2] = jp.nnz(); o[3] = hp.nnz();
        o[4] = P::nx; o[5] = P::nu; o[6] = P::nv; o[7] = P::np; o[8] = P::nb; o[9] = d.N;
        o[10] = d.step_block; o[11] = d.cblock;
    }
    void flags(int* o) override {
        o[0] = P::free_t0; o[1] = P::free_tf; o[2] = P::has_lagrange; o[3] = P::has_mayer; o[4] = P::is_max;
    }
    void structure(int which, int64_t* rows, int64_t* cols) override {
        const Pattern& p = which == 0 ? jp : hp;
        std::copy(p.rows.begin(), p.rows.end(), rows);
        std::copy(p.cols.begin(), p.cols.end(), cols);
    }
    void bounds(double* lvar, double* uvar, double* lcon, double* ucon) override {
        d.variable_bounds(lvar, uvar);
        d.constraint_bounds(lcon, ucon);
    }
    void x0(double* x) override { d.initial_guess(x); }
    void time_grid(const double* x, double* t) override {
        PlainAcc a{x};
        for (int i = 1; i <= d.N + 1; ++i) t[i - 1] = d.time_at(a, i);
    }
    double obj(const double* x) override { return d.objective(x); }
    void cons(const double* x, double* c) override { d.constraints(x, c); }
    void derivs(const double* x, const double* y, double sigma, double* grad, double* jvals, double* hvals,
                int64_t* dropped) override {
        typename Docp<P, SCH>::DerivOut o;
        o.jp = &jp; o.jvals = jvals;
        o.hp = &hp; o.hvals = hvals;
        o.grad = grad;
        d.derivatives(x, y, sigma, o);
        if (dropped) { dropped[0] = o.dropped_j; dropped[1] = o.dropped_h; }
    }
    int control_steps() const override { return Docp<P, SCH>::control_steps; }
    // T_control of build_OCP_solution (DOCP_data.jl:555-566): N * control_steps + 1 times
    void control_time(const double* x, double* tc) override {
        PlainAcc a{x};
        const int N = d.N, cs = Docp<P, SCH>::control_steps;
        int k = 0;
        for (int i = 1; i <= N; ++i) {
            const double ti = d.time_at(a, i), h_i = d.time_at(a, i + 1) - ti;
            for (int j = 1; j <= cs; ++j) tc[k++] = ti + (j - 1) * h_i / cs;
        }
        tc[k] = d.time_at(a, N + 1);
    }
    // getter() of src/ode/common.jl:7-104 + the normalisations of build_OCP_solution
    // (src/DOCP_data.jl:514-633): arrays are [node][component]; y may be NULL (then no duals)
    void unpack(const double* x, const double* y, double* time, double* state, double* control, double* variable,
                double* costate, double* path_dual, double* boundary_dual) override {
        PlainAcc a{x};
        const int N = d.N;
        std::vector<double> T(N + 1);
        for (int i = 1; i <= N + 1; ++i) T[i - 1] = d.time_at(a, i);
        for (int i = 1; i <= N + 1; ++i) {
            if (time) time[i - 1] = T[i - 1];
            if (state) for (int k = 0; k < P::nx; ++k) state[(i - 1) * P::nx + k] = x[d.state_off(i) + k];
            if (control && P::nu > 0) {
                // getter(:control), common.jl:82-99: N * control_steps + 1 rows, the last one u(tf)
                constexpr int cs = Docp<P, SCH>::control_steps;
                double u[P::nu > 0 ? P::nu : 1];
                if (cs == 1 || i == N + 1) {
                    d.get_control(a, i, u);
                    const int row = i == N + 1 ? N * cs : i - 1;
                    for (int k = 0; k < P::nu; ++k) control[row * P::nu + k] = u[k];
                } else {
                    for (int j = 1; j <= cs; ++j)
                        for (int k = 0; k < P::nu; ++k) control[((i - 1) * cs + (j - 1)) * P::nu + k] = x[d.control_off(i, j) + k];
                }
            }
        }
        if (variable) for (int k = 0; k < P::nv; ++k) variable[k] = x[d.v_off() + k];
        if (!y) return;
        for (int i = 1; i <= N; ++i)
            if (costate) for (int k = 0; k < P::nx; ++k) costate[(i - 1) * P::nx + k] = y[(i - 1) * d.cblock + k];
        if (path_dual)
            for (int i = 1; i <= N + 1; ++i) {
                const int off = i <= N ? (i - 1) * d.cblock + d.eqs_block : N * d.cblock;
                const double h = i <= N ? T[i] - T[i - 1] : T[N] - T[N - 1];   // DOCP_data.jl:596-602
                for (int r = 0; r < P::np; ++r) path_dual[(i - 1) * P::np + r] = y[off + r] / h;
            }
        if (boundary_dual) for (int r = 0; r < P::nb; ++r) boundary_dual[r] = y[d.ncon - P::nb + r];
    }
    // __initial_guess (DOCP_variables.jl:122-145, irk_stagewise.jl:302-335) from SAMPLED data: 0.1 everywhere,
    // then the variables, the states at the nodes and the controls (set_*_at_time_step!, common.jl:184-225;
    // the final control only where the scheme has one).  Stagewise schemes take one control per stage
    // (stage_control, [step][stage][k]) or, without it, the node control for every stage of the step.
    void pack(const double* state, const double* control, const double* variable, const double* stage_control,
              double* x) override {
        using ST = typename Docp<P, SCH>::ST;
        const int N = d.N;
        for (int k = 0; k < d.nvar; ++k) x[k] = 0.1;
        if (variable) for (int k = 0; k < P::nv; ++k) x[d.v_off() + k] = variable[k];
        for (int i = 1; i <= N + 1; ++i) {
            if (state) for (int k = 0; k < P::nx; ++k) x[d.state_off(i) + k] = state[(i - 1) * P::nx + k];
            if (P::nu == 0) continue;
            if (ST::stagewise) {
                if (i > N) continue;
                for (int j = 1; j <= ST::stage; ++j)
                    for (int k = 0; k < P::nu; ++k) {
                        if (stage_control) x[d.stagecontrol_off(i, j) + k] = stage_control[((i - 1) * ST::stage + (j - 1)) * P::nu + k];
                        else if (control) x[d.stagecontrol_off(i, j) + k] = control[(i - 1) * P::nu + k];
                    }
            } else if (control && (i <= N || ST::final_control)) {
                // the node's control for every control step of the time step (DOCP_variables.jl:138-140)
                for (int j = 1; j <= Docp<P, SCH>::control_steps; ++j)
                    for (int k = 0; k < P::nu; ++k) x[d.control_off(i, j) + k] = control[(i - 1) * P::nu + k];
            }
        }
    }
};

template <class P>
ModelBase* make_for_problem(int sch, int N, const double* tg, int ntg, const double* par, int npar, int pm) {
    switch (sch) {
        case TRAPEZE: return new Model<P, TRAPEZE>(N, tg, ntg, par, npar, pm);
        case MIDPOINT: return new Model<P, MIDPOINT>(N, tg, ntg, par, npar, pm);
        case GL1_CONST: return new Model<P, GL1_CONST>(N, tg, ntg, par, npar, pm);
        case GL2_CONST: return new Model<P, GL2_CONST>(N, tg, ntg, par, npar, pm);
        case GL3_CONST: return new Model<P, GL3_CONST>(N, tg, ntg, par, npar, pm);
        case GL2_STAGEWISE: return new Model<P, GL2_STAGEWISE>(N, tg, ntg, par, npar, pm);
        case GL3_STAGEWISE: return new Model<P, GL3_STAGEWISE>(N, tg, ntg, par, npar, pm);
        case EULER_EXPLICIT: return new Model<P, EULER_EXPLICIT>(N, tg, ntg, par, npar, pm);
        case EULER_IMPLICIT: return new Model<P, EULER_IMPLICIT>(N, tg, ntg, par, npar, pm);
        case MIDPOINT_CS2: return new Model<P, MIDPOINT_CS2>(N, tg, ntg, par, npar, pm);
        case MIDPOINT_CS3: return new Model<P, MIDPOINT_CS3>(N, tg, ntg, par, npar, pm);
        case MIDPOINT_CS10: return new Model<P, MIDPOINT_CS10>(N, tg, ntg, par, npar, pm);
    }
    return nullptr;
}

thread_local std::string g_err;

}  // namespace

extern "C" {

const char* orc_last_error() { return g_err.c_str(); }

void* orc_create(const char* problem, const char* scheme, int grid_size, const double* time_grid, int n_time_grid,
                 const double* params, int nparams, int pattern_mode) {
    try {
        int sch = scheme_from_name(scheme);
        if (sch < 0) { g_err = std::string("Unknown discretization method: ") + scheme; return nullptr; }
        std::string p = problem;
        ModelBase* m = nullptr;
#define ORC_P(T) if (p == T::name) m = make_for_problem<T>(sch, grid_size, time_grid, n_time_grid, params, nparams, pattern_mode);
        ORC_P(Goddard)
        ORC_P(GoddardAll)
        ORC_P(DoubleIntegratorMinEnergy)
        ORC_P(DoubleIntegratorMinTf)
        ORC_P(DoubleIntegratorFreeT0Tf)
        ORC_P(Beam)
        ORC_P(StagewiseScalar)
        ORC_P(BolzaFreeTf)
        ORC_P(VanDerPol)
        ORC_P(RobotArm)
        ORC_P(EstimateInitialCondition)
        ORC_P(EstimateRotationRate)
#undef ORC_P
        if (!m) g_err = std::string("unknown problem: ") + problem;
        return m;
    } catch (const std::exception& e) {
        g_err = e.what();
        return nullptr;
    }
}
void orc_destroy(void* h) { delete static_cast<ModelBase*>(h); }
void orc_dims(void* h, int64_t* out12) { static_cast<ModelBase*>(h)->dims(out12); }
void orc_flags(void* h, int* out5) { static_cast<ModelBase*>(h)->flags(out5); }
void orc_jac_structure(void* h, int64_t* rows, int64_t* cols) { static_cast<ModelBase*>(h)->structure(0, rows, cols); }
void orc_hess_structure(void* h, int64_t* rows, int64_t* cols) { static_cast<ModelBase*>(h)->structure(1, rows, cols); }
void orc_bounds(void* h, double* lvar, double* uvar, double* lcon, double* ucon) {
    static_cast<ModelBase*>(h)->bounds(lvar, uvar, lcon, ucon);
}
void orc_x0(void* h, double* x) { static_cast<ModelBase*>(h)->x0(x); }
void orc_time_grid(void* h, const double* x, double* t) { static_cast<ModelBase*>(h)->time_grid(x, t); }
double orc_obj(void* h, const double* x) { return static_cast<ModelBase*>(h)->obj(x); }
void orc_cons(void* h, const double* x, double* c) { static_cast<ModelBase*>(h)->cons(x, c); }
// any of grad / jvals / hvals may be NULL; y may be NULL (treated as 0)
void orc_derivs(void* h, const double* x, const double* y, double sigma, double* grad, double* jvals,
                double* hvals, int64_t* dropped2) {
    static_cast<ModelBase*>(h)->derivs(x, y, sigma, grad, jvals, hvals, dropped2);
}
void orc_unpack(void* h, const double* x, const double* y, double* time, double* state, double* control,
                double* variable, double* costate, double* path_dual, double* boundary_dual) {
    static_cast<ModelBase*>(h)->unpack(x, y, time, state, control, variable, costate, path_dual, boundary_dual);
}
int orc_control_steps(void* h) { return static_cast<ModelBase*>(h)->control_steps(); }
void orc_control_time(void* h, const double* x, double* tc) { static_cast<ModelBase*>(h)->control_time(x, tc); }

void orc_pack(void* h, const double* state, const double* control, const double* variable,
              const double* stage_control, double* x) {
    static_cast<ModelBase*>(h)->pack(state, control, variable, stage_control, x);
}
int orc_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
    return omp_get_max_threads();
#else
    (void)n;
    return 1;
#endif
}

}  // extern "C"
