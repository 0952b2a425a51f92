// TEST INFRASTRUCTURE ONLY -- part of the CPU oracle (see oracle/README.md).
//
// Optimal control problems used as oracle inputs, restated by hand as C++
// templates over the scalar type (double or oracle::HDual<N>).  These follow the
// callable signatures CTDirect consumes from CTModels:
//   dynamics(r, t, x, u, v), lagrange(t, x, u, v), mayer(x0, xf, v),
//   path(r, t, x, u, v), boundary(r, x0, xf, v), box triplets (lb, idx, ub).
// Sources: reference test/problems/goddard.jl:7-49,87-158,
//          test/problems/double_integrator.jl:4-19,40-56,79-99, test/problems/beam.jl:3-17,
//          test/ci/test_discretization_stagewise.jl:1-14,
//          test/problems/bolza.jl:3-18, test/problems/vanderpol.jl:3-19,
//          test/problems/autonomous_system.jl:6-90.
// The robot arm is NOT in the reference tree: it is a synthetic
// problem defined in this repository after the COPS robot-arm model.
#pragma once
#include <cmath>
#include <limits>
#include <string>
#include <vector>

#include "hdual.hpp"

namespace oracle {

constexpr double INF = std::numeric_limits<double>::infinity();

struct Box {  // (lb, 1-based index, ub) -- DOCP_variables.jl:88-98
    double lb;
    int idx;
    double ub;
};

struct ProblemMeta {
    std::vector<Box> state_box, control_box, variable_box;
    std::vector<double> path_lb, path_ub, bc_lb, bc_ub;
    std::vector<double> init_state, init_control, init_variable;  // empty => default 0.1
    std::vector<double> default_params;
};

// ---------------------------------------------------------------------------
// Goddard rocket, `@def` version: goddard.jl:18-49 (F0/F1 at l.7-15)
// params: Cd, beta, b, Tmax
// ---------------------------------------------------------------------------
struct Goddard {
    static constexpr const char* name = "goddard";
    static constexpr int nx = 3, nu = 1, nv = 1, np = 0, nb = 4, npar = 4;
    static constexpr bool free_t0 = false, free_tf = true, has_lagrange = false, has_mayer = true,
                          is_max = true;
    template <class T> static T t0(const T*, const double*) { return T(0.0); }
    template <class T> static T tf(const T* v, const double*) { return v[0]; }
    template <class T>
    static void dynamics(T* r, const T&, const T* x, const T* u, const T*, const double* p) {
        const double Cd = p[0], beta = p[1], b = p[2], Tmax = p[3];
        // F0(x) + u * F1(x)
        T D = Cd * sq(x[1]) * exp(-beta * (x[0] - 1.0));
        T f0_1 = x[1];
        T f0_2 = -D / x[2] - 1.0 / sq(x[0]);
        r[0] = f0_1 + u[0] * 0.0;
        r[1] = f0_2 + u[0] * (Tmax / x[2]);
        r[2] = 0.0 + u[0] * (-b * Tmax);
    }
    template <class T> static T lagrange(const T&, const T*, const T*, const T*, const double*) { return T(0.0); }
    template <class T> static T mayer(const T*, const T* xf, const T*, const double*) { return xf[0]; }
    template <class T> static void path(T*, const T&, const T*, const T*, const T*, const double*) {}
    template <class T>
    static void boundary(T* r, const T* x0, const T* xf, const T*, const double*) {
        r[0] = x0[0];
        r[1] = x0[1];
        r[2] = x0[2];
        r[3] = xf[2];
    }
    static ProblemMeta meta() {
        ProblemMeta m;
        const double r0 = 1, v0 = 0, m0 = 1, mf = 0.6, vmax = 0.1;
        m.state_box = {{r0, 1, r0 + 0.1}, {v0, 2, vmax}, {mf, 3, m0}};
        m.control_box = {{0.0, 1, 1.0}};
        m.variable_box = {{0.01, 1, INF}};
        m.bc_lb = {r0, v0, m0, mf};
        m.bc_ub = {r0, v0, m0, mf};
        m.init_state = {1.01, 0.05, 0.8};
        m.default_params = {310.0, 500.0, 2.0, 3.5};
        return m;
    }
};

// ---------------------------------------------------------------------------
// Goddard with all constraint kinds, functional version: goddard.jl:87-158
// ---------------------------------------------------------------------------
struct GoddardAll {
    static constexpr const char* name = "goddard_all";
    static constexpr int nx = 3, nu = 1, nv = 1, np = 3, nb = 4, npar = 4;
    static constexpr bool free_t0 = false, free_tf = true, has_lagrange = false, has_mayer = true,
                          is_max = true;
    template <class T> static T t0(const T*, const double*) { return T(0.0); }
    template <class T> static T tf(const T* v, const double*) { return v[0]; }
    template <class T>
    static void dynamics(T* r, const T&, const T* x, const T* u, const T*, const double* p) {
        const double Cd = p[0], beta = p[1], b = p[2], Tmax = p[3];
        r[0] = x[1];
        T D = Cd * sq(x[1]) * exp(-beta * (x[0] - 1.0));
        r[1] = -D / x[2] - 1.0 / sq(x[0]) + u[0] * Tmax / x[2];
        r[2] = -b * Tmax * u[0];
    }
    template <class T> static T lagrange(const T&, const T*, const T*, const T*, const double*) { return T(0.0); }
    template <class T> static T mayer(const T*, const T* xf, const T*, const double*) { return xf[0]; }
    template <class T>
    static void path(T* r, const T&, const T* x, const T* u, const T* v, const double*) {
        r[0] = x[1];
        r[1] = u[0];
        r[2] = x[0] + x[1] + x[2] + u[0] + v[0];
    }
    template <class T>
    static void boundary(T* r, const T* x0, const T* xf, const T*, const double*) {
        r[0] = x0[0];
        r[1] = x0[1];
        r[2] = x0[2];
        r[3] = xf[2];
    }
    static ProblemMeta meta() {
        ProblemMeta m;
        const double r0 = 1, v0 = 0, m0 = 1, mf = 0.6, vmax = 0.1;
        m.state_box = {{r0, 1, INF}, {v0, 2, INF}, {0.0, 3, m0}};
        m.control_box = {{0.0, 1, INF}};
        m.variable_box = {{0.01, 1, INF}};
        m.path_lb = {-INF, -INF, 0.0};
        m.path_ub = {vmax, 1.0, INF};
        m.bc_lb = {r0, v0, m0, mf};
        m.bc_ub = {r0, v0, m0, mf};
        m.init_state = {1.01, 0.05, 0.8};
        m.default_params = {310.0, 500.0, 2.0, 3.5};
        return m;
    }
};

// ---------------------------------------------------------------------------
// Double integrator, min energy, fixed T: double_integrator.jl:40-56
// params: T (final time, default 2)
// ---------------------------------------------------------------------------
struct DoubleIntegratorMinEnergy {
    static constexpr const char* name = "double_integrator_minenergy";
    static constexpr int nx = 2, nu = 1, nv = 0, np = 0, nb = 4, npar = 1;
    static constexpr bool free_t0 = false, free_tf = false, has_lagrange = true, has_mayer = false,
                          is_max = false;
    template <class T> static T t0(const T*, const double*) { return T(0.0); }
    template <class T> static T tf(const T*, const double* p) { return T(p[0]); }
    template <class T>
    static void dynamics(T* r, const T&, const T* x, const T* u, const T*, const double*) {
        r[0] = x[1];
        r[1] = u[0];
    }
    template <class T> static T lagrange(const T&, const T*, const T* u, const T*, const double*) { return sq(u[0]); }
    template <class T> static T mayer(const T*, const T*, const T*, const double*) { return T(0.0); }
    template <class T> static void path(T*, const T&, const T*, const T*, const T*, const double*) {}
    template <class T>
    static void boundary(T* r, const T* x0, const T* xf, const T*, const double*) {
        r[0] = x0[0];
        r[1] = x0[1];
        r[2] = xf[0];
        r[3] = xf[1];
    }
    static ProblemMeta meta() {
        ProblemMeta m;
        m.bc_lb = {0, 0, 1, 0};
        m.bc_ub = {0, 0, 1, 0};
        m.default_params = {2.0};
        return m;
    }
};

// ---------------------------------------------------------------------------
// Double integrator, min tf: double_integrator.jl:4-19
// ---------------------------------------------------------------------------
struct DoubleIntegratorMinTf {
    static constexpr const char* name = "double_integrator_mintf";
    static constexpr int nx = 2, nu = 1, nv = 1, np = 0, nb = 4, npar = 0;
    static constexpr bool free_t0 = false, free_tf = true, has_lagrange = false, has_mayer = true,
                          is_max = false;
    template <class T> static T t0(const T*, const double*) { return T(0.0); }
    template <class T> static T tf(const T* v, const double*) { return v[0]; }
    template <class T>
    static void dynamics(T* r, const T&, const T* x, const T* u, const T*, const double*) {
        r[0] = x[1];
        r[1] = u[0];
    }
    template <class T> static T lagrange(const T&, const T*, const T*, const T*, const double*) { return T(0.0); }
    template <class T> static T mayer(const T*, const T*, const T* v, const double*) { return v[0]; }
    template <class T> static void path(T*, const T&, const T*, const T*, const T*, const double*) {}
    template <class T>
    static void boundary(T* r, const T* x0, const T* xf, const T*, const double*) {
        r[0] = x0[0];
        r[1] = x0[1];
        r[2] = xf[0];
        r[3] = xf[1];
    }
    static ProblemMeta meta() {
        ProblemMeta m;
        m.control_box = {{-1.0, 1, 1.0}};
        m.variable_box = {{0.05, 1, INF}};
        m.bc_lb = {0, 0, 1, 0};
        m.bc_ub = {0, 0, 1, 0};
        return m;
    }
};

// ---------------------------------------------------------------------------
// Double integrator with free t0 and tf, max t0: double_integrator.jl:79-99
// (`0.01 <= tf - t0` depends on the variable only: a boundary constraint for CTModels, declared last)
// ---------------------------------------------------------------------------
struct DoubleIntegratorFreeT0Tf {
    static constexpr const char* name = "double_integrator_freet0tf";
    static constexpr int nx = 2, nu = 1, nv = 2, np = 0, nb = 5, npar = 0;
    static constexpr bool free_t0 = true, free_tf = true, has_lagrange = false, has_mayer = true,
                          is_max = true;
    template <class T> static T t0(const T* v, const double*) { return v[0]; }
    template <class T> static T tf(const T* v, const double*) { return v[1]; }
    template <class T>
    static void dynamics(T* r, const T&, const T* x, const T* u, const T*, const double*) {
        r[0] = x[1];
        r[1] = u[0];
    }
    template <class T> static T lagrange(const T&, const T*, const T*, const T*, const double*) { return T(0.0); }
    template <class T> static T mayer(const T*, const T*, const T* v, const double*) { return v[0]; }
    template <class T> static void path(T*, const T&, const T*, const T*, const T*, const double*) {}
    template <class T>
    static void boundary(T* r, const T* x0, const T* xf, const T* v, const double*) {
        r[0] = x0[0];
        r[1] = x0[1];
        r[2] = xf[0];
        r[3] = xf[1];
        r[4] = v[1] - v[0];
    }
    static ProblemMeta meta() {
        ProblemMeta m;
        m.control_box = {{-1.0, 1, 1.0}};
        m.variable_box = {{0.05, 1, 10.0}, {0.05, 2, 10.0}};
        m.bc_lb = {0, 0, 1, 0, 0.01};
        m.bc_ub = {0, 0, 1, 0, INF};
        return m;
    }
};

// ---------------------------------------------------------------------------
// Beam (bocop), Lagrange cost, state and control boxes: beam.jl:3-17
// ---------------------------------------------------------------------------
struct Beam {
    static constexpr const char* name = "beam";
    static constexpr int nx = 2, nu = 1, nv = 0, np = 0, nb = 4, npar = 0;
    static constexpr bool free_t0 = false, free_tf = false, has_lagrange = true, has_mayer = false,
                          is_max = false;
    template <class T> static T t0(const T*, const double*) { return T(0.0); }
    template <class T> static T tf(const T*, const double*) { return T(1.0); }
    template <class T>
    static void dynamics(T* r, const T&, const T* x, const T* u, const T*, const double*) {
        r[0] = x[1];
        r[1] = u[0];
    }
    template <class T> static T lagrange(const T&, const T*, const T* u, const T*, const double*) { return sq(u[0]); }
    template <class T> static T mayer(const T*, const T*, const T*, const double*) { return T(0.0); }
    template <class T> static void path(T*, const T&, const T*, const T*, const T*, const double*) {}
    template <class T>
    static void boundary(T* r, const T* x0, const T* xf, const T*, const double*) {
        r[0] = x0[0];
        r[1] = x0[1];
        r[2] = xf[0];
        r[3] = xf[1];
    }
    static ProblemMeta meta() {
        ProblemMeta m;
        m.state_box = {{0.0, 1, 0.1}};
        m.control_box = {{-10.0, 1, 10.0}};
        m.bc_lb = {0, 1, 0, -1};
        m.bc_ub = {0, 1, 0, -1};
        return m;
    }
};

// ---------------------------------------------------------------------------
// Scalar known-answer problem: test_discretization_stagewise.jl:1-14
// ---------------------------------------------------------------------------
struct StagewiseScalar {
    static constexpr const char* name = "stagewise_scalar";
    static constexpr int nx = 1, nu = 1, nv = 0, np = 0, nb = 2, npar = 0;
    static constexpr bool free_t0 = false, free_tf = false, has_lagrange = true, has_mayer = false,
                          is_max = false;
    template <class T> static T t0(const T*, const double*) { return T(0.0); }
    template <class T> static T tf(const T*, const double*) { return T(1.0); }
    template <class T>
    static void dynamics(T* r, const T&, const T*, const T* u, const T*, const double*) { r[0] = u[0]; }
    template <class T> static T lagrange(const T&, const T*, const T* u, const T*, const double*) { return sq(u[0]); }
    template <class T> static T mayer(const T*, const T*, const T*, const double*) { return T(0.0); }
    template <class T> static void path(T*, const T&, const T*, const T*, const T*, const double*) {}
    template <class T>
    static void boundary(T* r, const T* x0, const T* xf, const T*, const double*) {
        r[0] = x0[0];
        r[1] = xf[0];
    }
    static ProblemMeta meta() {
        ProblemMeta m;
        m.control_box = {{0.0, 1, 2.0}};
        m.bc_lb = {0, 1};
        m.bc_ub = {0, 1};
        return m;
    }
};

// ---------------------------------------------------------------------------
// Bolza problem with free tf entering the dynamics: bolza.jl:3-18
// ---------------------------------------------------------------------------
struct BolzaFreeTf {
    static constexpr const char* name = "bolza_freetf";
    static constexpr int nx = 1, nu = 1, nv = 1, np = 0, nb = 2, npar = 0;
    static constexpr bool free_t0 = false, free_tf = true, has_lagrange = true, has_mayer = true,
                          is_max = false;
    template <class T> static T t0(const T*, const double*) { return T(0.0); }
    template <class T> static T tf(const T* v, const double*) { return v[0]; }
    template <class T>
    static void dynamics(T* r, const T&, const T*, const T* u, const T* v, const double*) { r[0] = v[0] * u[0]; }
    template <class T> static T lagrange(const T&, const T*, const T* u, const T*, const double*) { return 0.5 * sq(u[0]); }
    template <class T> static T mayer(const T*, const T*, const T* v, const double*) { return v[0]; }
    template <class T> static void path(T*, const T&, const T*, const T*, const T*, const double*) {}
    template <class T>
    static void boundary(T* r, const T* x0, const T* xf, const T*, const double*) {
        r[0] = x0[0];
        r[1] = xf[0];
    }
    static ProblemMeta meta() {
        ProblemMeta m;
        m.state_box = {{0.0, 1, INF}};
        m.variable_box = {{0.1, 1, INF}};
        m.bc_lb = {0, 1};
        m.bc_ub = {0, 1};
        return m;
    }
};

// ---------------------------------------------------------------------------
// Van der Pol: vanderpol.jl:3-19  (params: omega, epsilon)
// ---------------------------------------------------------------------------
struct VanDerPol {
    static constexpr const char* name = "vanderpol";
    static constexpr int nx = 2, nu = 1, nv = 0, np = 0, nb = 2, npar = 2;
    static constexpr bool free_t0 = false, free_tf = false, has_lagrange = true, has_mayer = false,
                          is_max = false;
    template <class T> static T t0(const T*, const double*) { return T(0.0); }
    template <class T> static T tf(const T*, const double*) { return T(2.0); }
    template <class T>
    static void dynamics(T* r, const T&, const T* x, const T* u, const T*, const double* p) {
        const double omega = p[0], epsilon = p[1];
        r[0] = x[1];
        r[1] = epsilon * omega * (1.0 - sq(x[0])) * x[1] - (omega * omega) * x[0] + u[0];
    }
    template <class T>
    static T lagrange(const T&, const T* x, const T* u, const T*, const double*) {
        return 0.5 * (sq(x[0]) + sq(x[1]) + sq(u[0]));
    }
    template <class T> static T mayer(const T*, const T*, const T*, const double*) { return T(0.0); }
    template <class T> static void path(T*, const T&, const T*, const T*, const T*, const double*) {}
    template <class T>
    static void boundary(T* r, const T* x0, const T*, const T*, const double*) {
        r[0] = x0[0];
        r[1] = x0[1];
    }
    static ProblemMeta meta() {
        ProblemMeta m;
        m.bc_lb = {1, 0};
        m.bc_ub = {1, 0};
        m.default_params = {1.0, 1.0};
        return m;
    }
};

// ---------------------------------------------------------------------------
// Robot arm (SYNTHETIC, after COPS / OptimalControlProblems.jl `robot`; not in
// the reference tree).  x = (rho, rho', theta, theta', phi, phi'), u in [-1,1]^3,
// v = tf (min).  params: L
// ---------------------------------------------------------------------------
struct RobotArm {
    static constexpr const char* name = "robot";
    static constexpr int nx = 6, nu = 3, nv = 1, np = 0, nb = 12, npar = 1;
    static constexpr bool free_t0 = false, free_tf = true, has_lagrange = false, has_mayer = true,
                          is_max = false;
    template <class T> static T t0(const T*, const double*) { return T(0.0); }
    template <class T> static T tf(const T* v, const double*) { return v[0]; }
    template <class T>
    static void dynamics(T* r, const T&, const T* x, const T* u, const T*, const double* p) {
        const double L = p[0];
        T Iphi = (cube(L - x[0]) + cube(x[0])) / 3.0;
        T sphi = sin(x[4]);
        T Itheta = Iphi * sq(sphi);
        r[0] = x[1];
        r[1] = u[0] / L;
        r[2] = x[3];
        r[3] = u[1] / Itheta;
        r[4] = x[5];
        r[5] = u[2] / Iphi;
    }
    template <class T> static T lagrange(const T&, const T*, const T*, const T*, const double*) { return T(0.0); }
    template <class T> static T mayer(const T*, const T*, const T* v, const double*) { return v[0]; }
    template <class T> static void path(T*, const T&, const T*, const T*, const T*, const double*) {}
    template <class T>
    static void boundary(T* r, const T* x0, const T* xf, const T*, const double*) {
        for (int k = 0; k < 6; ++k) r[k] = x0[k];
        for (int k = 0; k < 6; ++k) r[6 + k] = xf[k];
    }
    static ProblemMeta meta() {
        ProblemMeta m;
        const double pi = 3.14159265358979323846;
        m.state_box = {{0.0, 1, 5.0}, {-pi, 3, pi}, {0.0, 5, pi}};
        m.control_box = {{-1.0, 1, 1.0}, {-1.0, 2, 1.0}, {-1.0, 3, 1.0}};
        m.variable_box = {{0.1, 1, INF}};
        m.bc_lb = {4.5, 0, 0, 0, pi / 4, 0, 4.5, 0, 2 * pi / 3, 0, pi / 4, 0};
        m.bc_ub = m.bc_lb;
        m.init_state = {4.5, 0.0, 1.0, 0.0, pi / 4, 0.0};
        m.init_control = {0.0, 0.0, 0.0};
        m.init_variable = {1.0};
        m.default_params = {5.0};
        return m;
    }
};

// ---------------------------------------------------------------------------
// Parameter-estimation problems WITHOUT control (NLP_u = 0): test/problems/autonomous_system.jl:6-90
// ---------------------------------------------------------------------------
struct EstimateInitialCondition {
    static constexpr const char* name = "estimate_initial_condition";
    static constexpr int nx = 2, nu = 0, nv = 2, np = 0, nb = 2, npar = 0;
    static constexpr bool free_t0 = false, free_tf = false, has_lagrange = false, has_mayer = true,
                          is_max = false;
    template <class T> static T t0(const T*, const double*) { return T(0.0); }
    template <class T> static T tf(const T*, const double*) { return T(3.14159265358979323846 / 2); }
    template <class T>
    static void dynamics(T* r, const T&, const T* x, const T*, const T*, const double*) {
        r[0] = -x[1];
        r[1] = x[0];
    }
    template <class T> static T lagrange(const T&, const T*, const T*, const T*, const double*) { return T(0.0); }
    template <class T> static T mayer(const T*, const T* xf, const T*, const double*) {
        return sq(xf[0] - 0.0) + sq(xf[1] - 1.0);
    }
    template <class T> static void path(T*, const T&, const T*, const T*, const T*, const double*) {}
    template <class T>
    static void boundary(T* r, const T* x0, const T*, const T* v, const double*) {
        r[0] = x0[0] - v[0];
        r[1] = x0[1] - v[1];
    }
    static ProblemMeta meta() {
        ProblemMeta m;
        m.bc_lb = {0, 0};
        m.bc_ub = {0, 0};
        return m;
    }
};

struct EstimateRotationRate {
    static constexpr const char* name = "estimate_rotation_rate";
    static constexpr int nx = 2, nu = 0, nv = 1, np = 0, nb = 2, npar = 0;
    static constexpr bool free_t0 = false, free_tf = false, has_lagrange = false, has_mayer = true,
                          is_max = false;
    template <class T> static T t0(const T*, const double*) { return T(0.0); }
    template <class T> static T tf(const T*, const double*) { return T(1.0); }
    template <class T>
    static void dynamics(T* r, const T&, const T* x, const T*, const T* v, const double*) {
        r[0] = v[0] * (-x[1]);
        r[1] = v[0] * x[0];
    }
    template <class T> static T lagrange(const T&, const T*, const T*, const T*, const double*) { return T(0.0); }
    template <class T> static T mayer(const T*, const T* xf, const T* v, const double*) {
        return sq(xf[0] - 0.0) + sq(xf[1] - 1.0) + 0.01 * sq(v[0]);
    }
    template <class T> static void path(T*, const T&, const T*, const T*, const T*, const double*) {}
    template <class T>
    static void boundary(T* r, const T* x0, const T*, const T*, const double*) {
        r[0] = x0[0] - 1.0;
        r[1] = x0[1] - 0.0;
    }
    static ProblemMeta meta() {
        ProblemMeta m;
        m.variable_box = {{0.0, 1, 10.0}};
        m.bc_lb = {0, 0};
        m.bc_ub = {0, 0};
        return m;
    }
};

}  // namespace oracle
