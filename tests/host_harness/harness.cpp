// TEST INFRASTRUCTURE ONLY.  Compiles ONE generated element header (csrc/generated/<E>.cuh) for the
// HOST with g++ (the CUDA qualifiers are defined away on the command line) and exposes its element
// functions through a C interface, so that the generator's derivative code can be checked against the
// oracle on a machine without a GPU.  The product never builds or loads this file.
#include <cmath>
#include <cstring>
using std::exp; using std::log; using std::sin; using std::cos; using std::sqrt;
#include CTD_ELEMENT_HEADER

using E = ctd_gen::CTD_ELEMENT;

struct Rec {
    double *c_, *jo_, *jn_, *jv_, *ho_, *hn_, *hv_, *go_, *gn_, *gv_;
    template <int R> void c(double v) { c_[R] = v; }
    template <int N> void jo(double v) { jo_[N] = v; }
    template <int N> void jn(double v) { jn_[N] = v; }
    template <int N> void jv(double v) { jv_[N] = v; }
    template <int N> void ho(double v) { ho_[N] = v; }
    template <int N> void hn(double v) { hn_[N] = v; }
    template <int K> void hv(double v) { hv_[K] += v; }
    template <int N> void go(double v) { go_[N] = v; }
    template <int N> void gn(double v) { gn_[N] = v; }
    template <int N> void gv(double v) { gv_[E::G_V(N, 2)] += v; }
};
struct RecEdge {
    double *c_, *j_, *h_, *g_;
    template <int R> void c(double v) { c_[R] = v; }
    template <int N> void j(double v) { j_[N] = v; }
    template <int N> void h(double v) { h_[N] = v; }
    template <int N> void g(double v) { g_[N] = v; }
};

extern "C" {
// counts: NZ, CB, NJ_OWN, NJ_NEXT, NJ_V, NH_OWN, NH_NEXT, NVV, NG_OWN, NG_NEXT, NV, FP_NLOC, NP, NFP_J, NFP_H,
//         BC_NLOC, NB, NBC_J, NBC_H, NBC_G, NOWN, NNEXT, BLK, NX, NPAR, T0_KIND, TF_KIND, T0_IDX, TF_IDX, EQS
void elem_dims(int* o) {
    const int v[30] = {E::NZ, E::CB, E::NJ_OWN, E::NJ_NEXT, E::NJ_V, E::NH_OWN, E::NH_NEXT, E::NV * (E::NV + 1) / 2,
                       E::NG_OWN, E::NG_NEXT, E::NV, E::FP_NLOC, E::NP, E::NFP_J, E::NFP_H, E::BC_NLOC, E::NB,
                       E::NBC_J, E::NBC_H, E::NBC_G, E::NOWN, E::NNEXT, E::BLK, E::NX, E::NPAR, E::T0_KIND,
                       E::TF_KIND, E::T0_IDX, E::TF_IDX, E::EQS};
    std::memcpy(o, v, sizeof(v));
}
void elem_consts(double* o) { o[0] = E::T0_CONST; o[1] = E::TF_CONST; }
// table ids: 0 J_OWN 1 J_NEXT 2 J_V 3 H_OWN 4 H_NEXT 5 FP_J 6 FP_H 7 BC_J 8 BC_H 9 G_OWN 10 G_NEXT 11 G_V 12 BC_G
int elem_table(int id, int* out) {
    int n = 0;
#define T_(ID, NAME) if (id == ID) { n = E::N##NAME; for (int i = 0; i < n; ++i) for (int k = 0; k < 3; ++k) out[3 * i + k] = E::NAME(i, k); }
    T_(0, J_OWN) T_(1, J_NEXT) T_(2, J_V) T_(3, H_OWN) T_(4, H_NEXT) T_(5, FP_J) T_(6, FP_H) T_(7, BC_J) T_(8, BC_H)
    T_(9, G_OWN) T_(10, G_NEXT) T_(11, G_V) T_(12, BC_G)
#undef T_
    return n;
}
static ctd::TimeArgs<E::NV> make_T(const double* t) {   // ti, tp, hh, then dti[NV], dtp[NV], dh[NV]
    ctd::TimeArgs<E::NV> T;
    T.ti = t[0]; T.tp = t[1]; T.hh = t[2];
    for (int k = 0; k < E::NV; ++k) { T.dti[k] = t[3 + k]; T.dtp[k] = t[3 + E::NV + k]; T.dh[k] = t[3 + 2 * E::NV + k]; }
    return T;
}
double elem_step(const double* z, const double* t, const double* p, const double* lam, double sg, double* c,
                 double* jo, double* jn, double* jv, double* ho, double* hn, double* hv, double* go, double* gn,
                 double* gv) {
    Rec s{c, jo, jn, jv, ho, hn, hv, go, gn, gv};
    const ctd::TimeArgs<E::NV> T = make_T(t);
    E::template step<true, true, true>(z, T, p, lam, sg, s);
    E::grad_step(z, T, p, s);
    return E::cost_step(z, T, p);
}
void elem_fpath(const double* z, const double* t, const double* p, const double* lam, double* c, double* j, double* h) {
    RecEdge s{c, j, h, nullptr};
    E::template fpath<true, true, true>(z, make_T(t), p, lam, s);
}
double elem_bc(const double* z, const double* p, const double* lam, double sg, double* c, double* j, double* h, double* g) {
    RecEdge s{c, j, h, g};
    E::template bc<true, true, true>(z, p, lam, sg, s);
    E::grad_bc(z, p, s);
    return E::mayer(z, p);
}
}
