// TEST INFRASTRUCTURE ONLY -- part of the CPU oracle.
//
// Sparsity tracer: an operator-overloading scalar that records which local variables a quantity depends
// on (first order: a set of indices; second order: a set of index pairs), without looking at values.
// It follows the operator classification of SparseConnectivityTracer's global tracers [3P, v1.2.1 named in
// the reference's Project.toml:36], which is what ADNLPModels' `:optimized` backend runs over the
// reference's closures (src/collocation.jl:131-134) to DETECT the Jacobian / Hessian patterns:
//   a + b, a - b      first order in both, no second-order term
//   a * b             adds the cross pairs  grad(a) x grad(b)
//   a / b             adds  grad(a) x grad(b)  and  grad(b) x grad(b)
//   c * a, a * c, a / c, a + c (c a plain number)   linear: dependencies kept even when c == 0
//   exp, log, sin, cos, sqrt, 1/a, a^n             adds  grad(a) x grad(a)
// The reference pins the resulting counts for Goddard / midpoint / N = 250: nnzj 4504, nnzh 5259
// (test/ci/test_modeler_solver.jl:32).
#pragma once
#include <cstdint>

namespace oracle {

template <int N>
struct Tracer {
    static_assert(N <= 64, "tracer supports at most 64 local variables");
    uint64_t g = 0;                 // first-order dependencies
    uint64_t h[N > 0 ? N : 1];      // h[a] bit b: the pair (a, b) may have a nonzero second derivative
    Tracer() { clear(); }
    Tracer(double) { clear(); }
    void clear() {
        g = 0;
        for (int a = 0; a < N; ++a) h[a] = 0;
    }
    static Tracer variable(double, int dir) {
        Tracer r;
        r.g = uint64_t(1) << dir;
        return r;
    }
    bool dep(int a) const { return (g >> a) & 1; }
    bool pair(int a, int b) const { return (h[a] >> b) & 1; }
    void add_cross(uint64_t ga, uint64_t gb) {  // symmetric closure of ga x gb
        for (int a = 0; a < N; ++a) {
            if ((ga >> a) & 1) h[a] |= gb;
            if ((gb >> a) & 1) h[a] |= ga;
        }
    }
};

template <int N>
inline Tracer<N> join(const Tracer<N>& x, const Tracer<N>& y) {
    Tracer<N> r;
    r.g = x.g | y.g;
    for (int a = 0; a < N; ++a) r.h[a] = x.h[a] | y.h[a];
    return r;
}
template <int N>
inline Tracer<N> nonlinear(const Tracer<N>& x) {
    Tracer<N> r = x;
    r.add_cross(x.g, x.g);
    return r;
}
template <int N> inline Tracer<N> operator+(const Tracer<N>& x, const Tracer<N>& y) { return join(x, y); }
template <int N> inline Tracer<N> operator-(const Tracer<N>& x, const Tracer<N>& y) { return join(x, y); }
template <int N> inline Tracer<N> operator-(const Tracer<N>& x) { return x; }
template <int N>
inline Tracer<N> operator*(const Tracer<N>& x, const Tracer<N>& y) {
    Tracer<N> r = join(x, y);
    r.add_cross(x.g, y.g);
    return r;
}
template <int N>
inline Tracer<N> operator/(const Tracer<N>& x, const Tracer<N>& y) {
    Tracer<N> r = join(x, y);
    r.add_cross(x.g, y.g);
    r.add_cross(y.g, y.g);
    return r;
}
// mixed with plain numbers: linear (the value of the number is never inspected)
template <int N> inline Tracer<N> operator+(const Tracer<N>& x, double) { return x; }
template <int N> inline Tracer<N> operator+(double, const Tracer<N>& x) { return x; }
template <int N> inline Tracer<N> operator-(const Tracer<N>& x, double) { return x; }
template <int N> inline Tracer<N> operator-(double, const Tracer<N>& x) { return x; }
template <int N> inline Tracer<N> operator*(const Tracer<N>& x, double) { return x; }
template <int N> inline Tracer<N> operator*(double, const Tracer<N>& x) { return x; }
template <int N> inline Tracer<N> operator/(const Tracer<N>& x, double) { return x; }
template <int N> inline Tracer<N> operator/(double, const Tracer<N>& x) { return nonlinear(x); }
template <int N> inline Tracer<N>& operator+=(Tracer<N>& x, const Tracer<N>& y) { x = x + y; return x; }
template <int N> inline Tracer<N>& operator-=(Tracer<N>& x, const Tracer<N>& y) { x = x - y; return x; }

template <int N> inline Tracer<N> exp(const Tracer<N>& x) { return nonlinear(x); }
template <int N> inline Tracer<N> log(const Tracer<N>& x) { return nonlinear(x); }
template <int N> inline Tracer<N> sin(const Tracer<N>& x) { return nonlinear(x); }
template <int N> inline Tracer<N> cos(const Tracer<N>& x) { return nonlinear(x); }
template <int N> inline Tracer<N> sqrt(const Tracer<N>& x) { return nonlinear(x); }
template <int N> inline Tracer<N> sq(const Tracer<N>& x) { return nonlinear(x); }     // x^2
template <int N> inline Tracer<N> cube(const Tracer<N>& x) { return nonlinear(x); }   // x^3

}  // namespace oracle
