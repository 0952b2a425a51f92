// TEST INFRASTRUCTURE ONLY -- part of the CPU oracle (see oracle/README.md).
//
// Second-order forward-mode automatic differentiation ("hyper-dual" numbers
// carrying a value, a gradient and a packed symmetric Hessian with respect to
// N seeded directions).  The reference obtains its derivatives by
// operator-overloading AD inside ADNLPModels (src/collocation.jl:105-149 of the
// reference, [3P] ForwardDiff / ReverseDiff); this file is the oracle's
// equivalent so that the oracle shares NO derivative code with the generated
// analytic CUDA kernels.
#pragma once
#include <cmath>

namespace oracle {

template <int N>
struct HDual {
    static constexpr int NH = N * (N + 1) / 2;
    double v;
    double g[N > 0 ? N : 1];
    double h[NH > 0 ? NH : 1];  // packed lower triangle: (a,b), a>=b at a*(a+1)/2+b

    HDual() : v(0.0) { zero_d(); }
    HDual(double c) : v(c) { zero_d(); }
    static HDual variable(double val, int dir) {
        HDual r(val);
        r.g[dir] = 1.0;
        return r;
    }
    void zero_d() {
        for (int a = 0; a < N; ++a) g[a] = 0.0;
        for (int a = 0; a < NH; ++a) h[a] = 0.0;
    }
    static int hidx(int a, int b) { return a >= b ? a * (a + 1) / 2 + b : b * (b + 1) / 2 + a; }
    double hess(int a, int b) const { return h[hidx(a, b)]; }
};

// generic unary chain rule: y = f(u), with f0=f(u.v), f1=f'(u.v), f2=f''(u.v)
template <int N>
inline HDual<N> chain(const HDual<N>& u, double f0, double f1, double f2) {
    HDual<N> r;
    r.v = f0;
    for (int a = 0; a < N; ++a) r.g[a] = f1 * u.g[a];
    int k = 0;
    for (int a = 0; a < N; ++a)
        for (int b = 0; b <= a; ++b, ++k) r.h[k] = f1 * u.h[k] + f2 * (u.g[a] * u.g[b]);
    return r;
}

template <int N>
inline HDual<N> operator+(const HDual<N>& x, const HDual<N>& y) {
    HDual<N> r;
    r.v = x.v + y.v;
    for (int a = 0; a < N; ++a) r.g[a] = x.g[a] + y.g[a];
    for (int a = 0; a < HDual<N>::NH; ++a) r.h[a] = x.h[a] + y.h[a];
    return r;
}
template <int N>
inline HDual<N> operator-(const HDual<N>& x, const HDual<N>& y) {
    HDual<N> r;
    r.v = x.v - y.v;
    for (int a = 0; a < N; ++a) r.g[a] = x.g[a] - y.g[a];
    for (int a = 0; a < HDual<N>::NH; ++a) r.h[a] = x.h[a] - y.h[a];
    return r;
}
template <int N>
inline HDual<N> operator-(const HDual<N>& x) {
    HDual<N> r;
    r.v = -x.v;
    for (int a = 0; a < N; ++a) r.g[a] = -x.g[a];
    for (int a = 0; a < HDual<N>::NH; ++a) r.h[a] = -x.h[a];
    return r;
}
template <int N>
inline HDual<N> operator*(const HDual<N>& x, const HDual<N>& y) {
    HDual<N> r;
    r.v = x.v * y.v;
    for (int a = 0; a < N; ++a) r.g[a] = x.g[a] * y.v + x.v * y.g[a];
    int k = 0;
    for (int a = 0; a < N; ++a)
        for (int b = 0; b <= a; ++b, ++k)
            r.h[k] = x.h[k] * y.v + x.v * y.h[k] + (x.g[a] * y.g[b] + x.g[b] * y.g[a]);
    return r;
}
template <int N>
inline HDual<N> inv(const HDual<N>& u) {
    const double i1 = 1.0 / u.v;
    return chain(u, i1, -i1 * i1, 2.0 * i1 * i1 * i1);
}
template <int N>
inline HDual<N> operator/(const HDual<N>& x, const HDual<N>& y) { return x * inv(y); }

// mixed with double
template <int N> inline HDual<N> operator+(const HDual<N>& x, double c) { HDual<N> r = x; r.v += c; return r; }
template <int N> inline HDual<N> operator+(double c, const HDual<N>& x) { return x + c; }
template <int N> inline HDual<N> operator-(const HDual<N>& x, double c) { HDual<N> r = x; r.v -= c; return r; }
template <int N> inline HDual<N> operator-(double c, const HDual<N>& x) { return (-x) + c; }
template <int N>
inline HDual<N> operator*(const HDual<N>& x, double c) {
    HDual<N> r;
    r.v = x.v * c;
    for (int a = 0; a < N; ++a) r.g[a] = x.g[a] * c;
    for (int a = 0; a < HDual<N>::NH; ++a) r.h[a] = x.h[a] * c;
    return r;
}
template <int N> inline HDual<N> operator*(double c, const HDual<N>& x) { return x * c; }
template <int N> inline HDual<N> operator/(const HDual<N>& x, double c) { return x * (1.0 / c); }
template <int N> inline HDual<N> operator/(double c, const HDual<N>& x) { return inv(x) * c; }
template <int N> inline HDual<N>& operator+=(HDual<N>& x, const HDual<N>& y) { x = x + y; return x; }
template <int N> inline HDual<N>& operator-=(HDual<N>& x, const HDual<N>& y) { x = x - y; return x; }

template <int N> inline HDual<N> exp(const HDual<N>& u) { const double e = std::exp(u.v); return chain(u, e, e, e); }
template <int N> inline HDual<N> log(const HDual<N>& u) { return chain(u, std::log(u.v), 1.0 / u.v, -1.0 / (u.v * u.v)); }
template <int N> inline HDual<N> sin(const HDual<N>& u) { const double s = std::sin(u.v), c = std::cos(u.v); return chain(u, s, c, -s); }
template <int N> inline HDual<N> cos(const HDual<N>& u) { const double s = std::sin(u.v), c = std::cos(u.v); return chain(u, c, -s, -c); }
template <int N> inline HDual<N> sqrt(const HDual<N>& u) { const double s = std::sqrt(u.v); return chain(u, s, 0.5 / s, -0.25 / (s * u.v)); }
template <int N> inline HDual<N> sq(const HDual<N>& u) { return u * u; }
template <int N> inline HDual<N> cube(const HDual<N>& u) { return u * u * u; }

// double overloads so that templated problem code can call the same names
inline double sq(double u) { return u * u; }
inline double cube(double u) { return u * u * u; }
using std::exp;
using std::log;
using std::sin;
using std::cos;
using std::sqrt;

inline double value_of(double x) { return x; }
template <int N> inline double value_of(const HDual<N>& x) { return x.v; }

}  // namespace oracle
