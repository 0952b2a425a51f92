// extern "C" entry points declared in include/ctdirect_b200.h.
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <exception>
#include <string>

#include "../../include/ctdirect_b200.h"
#include "ctd_model.hpp"

using namespace ctd;

static_assert(CTD_COMM_SLOTS == kCommSlots && CTD_NCCL_ID_BYTES == kNcclIdBytes, "header and model disagree");

// Per-handle call statistics (the NLPModels `Counters` of the reference's models): one count per
// callback kind, incremented by every entry point that produces that output, fused or not.
struct CallStats {
    std::atomic<int64_t> n[CTD_N_COUNTERS];
    std::atomic<int64_t> host_ns;   // wall time spent inside the blocking host-buffer calls
    CallStats() { reset(); }
    void reset() {
        for (auto& c : n) c.store(0, std::memory_order_relaxed);
        host_ns.store(0, std::memory_order_relaxed);
    }
    void count(int what, bool host, bool batched_norms = false) {
        if (what & CTD_OBJ) n[CTD_NEVAL_OBJ]++;
        if (what & CTD_GRAD) n[CTD_NEVAL_GRAD]++;
        if (what & CTD_CONS) n[CTD_NEVAL_CONS]++;
        if (what & CTD_JAC) n[CTD_NEVAL_JAC]++;
        if (what & CTD_HESS) n[CTD_NEVAL_HESS]++;
        n[CTD_NCALL_EVAL]++;
        if (host) n[CTD_NCALL_HOST]++;
        if (batched_norms) n[CTD_NCALL_NORMS]++;
    }
};

struct ctd_model_s {
    Model* m;
    CallStats stats;
};

namespace {
thread_local std::string g_err;

int fail(int code, const std::string& msg) {
    g_err = msg;
    return code == 0 ? -1 : code;
}

template <class Fn>
int guarded(Fn&& fn) {
    try {
        fn();
        return 0;
    } catch (const CtdError& e) {
        return fail(e.code, e.msg);
    } catch (const std::exception& e) {
        return fail(-1, e.what());
    } catch (...) {
        return fail(-1, "unknown error");
    }
}

int check(ctd_handle h) {
    if (!h || !h->m) return fail(-2, "invalid handle");
    return 0;
}
}  // namespace

extern "C" {

const char* ctd_last_error(void) { return g_err.c_str(); }

int64_t ctd_list_models(char* buf, int64_t buflen) {
    std::string s;
    for (const Registration& r : registry())
        s += r.info.problem + " " + r.info.scheme + " " +
             (r.info.pattern_mode == 2 ? "optimized" : r.info.pattern_mode ? "corrected" : "reference") + "\n";
    if (buf && buflen > 0) {
        const size_t n = std::min<size_t>(s.size(), (size_t)buflen - 1);
        std::memcpy(buf, s.data(), n);
        buf[n] = 0;
    }
    return (int64_t)s.size() + 1;
}

int ctd_create(const ctd_desc* d, ctd_handle* out) {
    if (!d || !out || !d->problem || !d->scheme) return fail(-2, "ctd_create: NULL argument");
    *out = nullptr;
    return guarded([&] {
        const Registration* reg = nullptr;
        bool scheme_known = false, problem_known = false;
        // scheme aliases of the reference's switch (src/DOCP_data.jl:315-320)
        std::string scheme = d->scheme;
        if (!scheme.empty() && scheme[0] == ':') scheme = scheme.substr(1);
        if (scheme == "euler_explicit" || scheme == "euler_forward") scheme = "euler";
        if (scheme == "euler_backward") scheme = "euler_implicit";
        int pattern_mode = d->pattern_mode;
        if (d->control_steps > 1) {   // DirectShooting: src/direct_shooting.jl:55-71
            if (scheme != "midpoint")
                throw CtdError{-5, "control_steps > 1 is compiled for the midpoint scheme only (direct_shooting.jl:16)"};
            scheme = "midpoint_cs" + std::to_string(d->control_steps);
            pattern_mode = 2;
        } else if (d->control_steps < 0) {
            throw CtdError{-2, "control_steps must be >= 0"};
        }
        for (const Registration& r : registry()) {
            if (r.info.scheme == scheme) scheme_known = true;
            if (r.info.problem == d->problem) problem_known = true;
            if (r.info.problem == d->problem && r.info.scheme == scheme) {
                if (r.info.pattern_mode == pattern_mode) reg = &r;
                // a pair without quirks is only compiled in reference mode: both manual modes coincide
                else if (!reg && r.info.pattern_mode == 0 && pattern_mode != 2) reg = &r;
            }
        }
        if (!reg) {
            if (!scheme_known && d->control_steps > 1)
                throw CtdError{-5, "control_steps = " + std::to_string(d->control_steps) + " is not compiled in (2, 3, 10 are)"};
            if (!scheme_known)  // same wording as the reference's error, DOCP_data.jl:343-348
                throw CtdError{-4, std::string("Unknown discretization method: ") + d->scheme};
            if (!problem_known) throw CtdError{-5, std::string("unknown problem: ") + d->problem};
            throw CtdError{-5, "this (problem, scheme) pair is not compiled in"};
        }
        if (d->pattern_mode < 0 || d->pattern_mode > 2) throw CtdError{-2, "pattern_mode must be 0, 1 or 2"};
        Model* m = new Model(*reg, d->grid_size, d->time_grid, d->n_time_grid, d->params, d->n_params,
                             d->batch < 1 ? 1 : d->batch, d->device);
        *out = new ctd_model_s;
        (*out)->m = m;
    });
}

int ctd_destroy(ctd_handle h) {
    if (!h) return 0;
    delete h->m;
    delete h;
    return 0;
}

int ctd_dims(ctd_handle h, int64_t* nvar, int64_t* ncon, int64_t* nnzj, int64_t* nnzh) {
    if (int e = check(h)) return e;
    if (nvar) *nvar = h->m->nvar();
    if (ncon) *ncon = h->m->ncon();
    if (nnzj) *nnzj = h->m->nnzj();
    if (nnzh) *nnzh = h->m->nnzh();
    return 0;
}

int ctd_info(ctd_handle h, int64_t out[16]) {
    if (int e = check(h)) return e;
    const ElemInfo& f = h->m->info();
    const int64_t v[16] = {f.nx, f.nu, f.nv, f.np, f.nb, h->m->N(), f.blk, f.cb, h->m->batch(),
                           f.free_t0, f.free_tf, f.lagrange, f.mayer, f.is_max, f.npar, f.stage};
    std::memcpy(out, v, sizeof(v));
    return 0;
}

int ctd_bounds(ctd_handle h, double* lvar, double* uvar, double* lcon, double* ucon) {
    if (int e = check(h)) return e;
    return guarded([&] {
        if (lvar && uvar) h->m->variable_bounds(lvar, uvar);
        if (lcon && ucon) h->m->constraint_bounds(lcon, ucon);
    });
}

int ctd_x0(ctd_handle h, double* x0) {
    if (int e = check(h)) return e;
    return guarded([&] { h->m->initial_guess(x0); });
}

int ctd_time_grid(ctd_handle h, double* s) {
    if (int e = check(h)) return e;
    const std::vector<double>& g = h->m->sgrid();
    std::memcpy(s, g.data(), sizeof(double) * g.size());
    return 0;
}

int ctd_jac_structure(ctd_handle h, int64_t* rows, int64_t* cols) {
    if (int e = check(h)) return e;
    return guarded([&] { h->m->structure(h->m->jac(), rows, cols); });
}

int ctd_hess_structure(ctd_handle h, int64_t* rows, int64_t* cols) {
    if (int e = check(h)) return e;
    return guarded([&] { h->m->structure(h->m->hess(), rows, cols); });
}

int ctd_eval(ctd_handle h, int what, const double* x, const double* y, double obj_weight, double* f, double* g,
             double* c, double* jvals, double* hvals) {
    if (int e = check(h)) return e;
    const auto t0 = std::chrono::steady_clock::now();
    const int rc = guarded([&] { h->m->eval_host(what, x, y, obj_weight, f, g, c, jvals, hvals); });
    h->stats.count(what, true);
    h->stats.host_ns += std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t0).count();
    return rc;
}

int ctd_eval_async(ctd_handle h, int what, const double* x, const double* y, double obj_weight, double* f, double* g,
                   double* c, double* jvals, double* hvals) {
    if (int e = check(h)) return e;
    h->stats.count(what, true);
    return guarded([&] { h->m->eval_host(what, x, y, obj_weight, f, g, c, jvals, hvals, false); });
}

int ctd_eval_dev(ctd_handle h, int what, const double* x, const double* y, double obj_weight, double* f,
                 double* g, double* c, double* jvals, double* hvals) {
    if (int e = check(h)) return e;
    h->stats.count(what, false);
    return guarded([&] { h->m->eval_dev(what, x, y, obj_weight, f, g, c, jvals, hvals); });
}

int ctd_eval_norms_dev(ctd_handle h, int what, const double* x, const double* y, double obj_weight, double* f,
                       double* g, double* c, double* jvals, double* hvals, double* viol) {
    if (int e = check(h)) return e;
    h->stats.count(what, false, true);
    return guarded([&] { h->m->eval_dev(what, x, y, obj_weight, f, g, c, jvals, hvals, viol); });
}

int ctd_obj(ctd_handle h, const double* x, double* f) {
    return ctd_eval(h, CTD_OBJ, x, nullptr, 1.0, f, nullptr, nullptr, nullptr, nullptr);
}
int ctd_grad(ctd_handle h, const double* x, double* g) {
    return ctd_eval(h, CTD_GRAD | CTD_OBJ, x, nullptr, 1.0, nullptr, g, nullptr, nullptr, nullptr);
}
int ctd_cons(ctd_handle h, const double* x, double* c) {
    return ctd_eval(h, CTD_CONS, x, nullptr, 1.0, nullptr, nullptr, c, nullptr, nullptr);
}
int ctd_jac_coord(ctd_handle h, const double* x, double* vals) {
    return ctd_eval(h, CTD_JAC, x, nullptr, 1.0, nullptr, nullptr, nullptr, vals, nullptr);
}
int ctd_hess_coord(ctd_handle h, const double* x, const double* y, double obj_weight, double* vals) {
    return ctd_eval(h, CTD_HESS, x, y, obj_weight, nullptr, nullptr, nullptr, nullptr, vals);
}

int ctd_norms_dev(ctd_handle h, const double* x, double* obj, double* viol) {
    if (int e = check(h)) return e;
    return guarded([&] { h->m->norms_dev(x, obj, viol); });
}

int ctd_violation_dev(ctd_handle h, const double* c, double* viol) {
    if (int e = check(h)) return e;
    return guarded([&] { h->m->violation_dev(c, viol); });
}

int ctd_norms(ctd_handle h, const double* x, double* obj, double* viol) {
    if (int e = check(h)) return e;
    return guarded([&] { h->m->norms_host(x, obj, viol); });
}

int ctd_unpack(ctd_handle h, const double* x, const double* y, const ctd_solution* o) {
    if (int e = check(h)) return e;
    if (!o) return fail(-2, "ctd_unpack: NULL output descriptor");
    return guarded([&] {
        h->m->unpack_host(x, y, o->time, o->state, o->control, o->variable, o->costate, o->path_dual, o->boundary_dual);
    });
}

int ctd_unpack_dev(ctd_handle h, const double* x, const double* y, const ctd_solution* o) {
    if (int e = check(h)) return e;
    if (!o) return fail(-2, "ctd_unpack_dev: NULL output descriptor");
    return guarded([&] {
        h->m->unpack_dev(x, y, o->time, o->state, o->control, o->variable, o->costate, o->path_dual, o->boundary_dual);
    });
}

int32_t ctd_control_steps(ctd_handle h) {
    if (check(h)) return -1;
    return h->m->info().control_steps;
}

int ctd_control_time(ctd_handle h, const double* time, double* control_time) {
    if (int e = check(h)) return e;
    if (!time || !control_time) return fail(-2, "ctd_control_time: NULL argument");
    const int64_t N = h->m->N(), cs = h->m->info().control_steps, B = h->m->batch();
    for (int64_t b = 0; b < B; ++b) {
        const double* T = time + b * (N + 1);
        double* tc = control_time + b * (N * cs + 1);
        int64_t k = 0;
        for (int64_t i = 0; i < N; ++i) {
            const double h_i = T[i + 1] - T[i];
            for (int64_t j = 0; j < cs; ++j) tc[k++] = T[i] + (double)j * h_i / (double)cs;   // DOCP_data.jl:562
        }
        tc[k] = T[N];
    }
    return 0;
}

int ctd_pack(ctd_handle h, const double* state, const double* control, const double* variable,
             const double* stage_control, double* x) {
    if (int e = check(h)) return e;
    return guarded([&] { h->m->pack_host(state, control, variable, stage_control, x); });
}

int ctd_pack_dev(ctd_handle h, const double* state, const double* control, const double* variable,
                 const double* stage_control, double* x) {
    if (int e = check(h)) return e;
    return guarded([&] { h->m->pack_dev(state, control, variable, stage_control, x); });
}

int ctd_set_params(ctd_handle h, const double* params, int32_t n) {
    if (int e = check(h)) return e;
    return guarded([&] { h->m->set_params(params, n); });
}

int ctd_sync(ctd_handle h) {
    if (int e = check(h)) return e;
    return guarded([&] { h->m->sync(); });
}

void* ctd_stream(ctd_handle h) {
    if (check(h)) return nullptr;
    void* s = nullptr;
    guarded([&] { s = (void*)h->m->stream(); });
    return s;
}

int ctd_set_stream(ctd_handle h, void* stream) {
    if (int e = check(h)) return e;
    return guarded([&] { h->m->set_stream((cudaStream_t)stream); });
}

int64_t ctd_launch_count(ctd_handle h) {
    if (check(h)) return -1;
    return h->m->launches();
}

int ctd_counters(ctd_handle h, int64_t out[CTD_N_COUNTERS]) {
    if (int e = check(h)) return e;
    if (!out) return fail(-2, "ctd_counters: NULL argument");
    for (int i = 0; i < CTD_N_COUNTERS; ++i) out[i] = h->stats.n[i].load(std::memory_order_relaxed);
    return 0;
}

int ctd_reset_counters(ctd_handle h) {
    if (int e = check(h)) return e;
    h->stats.reset();
    return 0;
}

int64_t ctd_stats_json(ctd_handle h, char* buf, int64_t buflen) {
    if (check(h)) return -1;
    const ElemInfo& f = h->m->info();
    const CallStats& st = h->stats;
    char tmp[1024];
    const int n = std::snprintf(
        tmp, sizeof(tmp),
        "{\"problem\": \"%s\", \"scheme\": \"%s\", \"grid_size\": %lld, \"batch\": %lld, \"nvar\": %lld, \"ncon\": %lld, "
        "\"nnzj\": %lld, \"nnzh\": %lld, \"neval_obj\": %lld, \"neval_grad\": %lld, \"neval_cons\": %lld, "
        "\"neval_jac\": %lld, \"neval_hess\": %lld, \"calls\": %lld, \"host_calls\": %lld, \"norm_calls\": %lld, "
        "\"kernel_launches\": %lld, \"host_call_seconds\": %.9f}",
        f.problem.c_str(), f.scheme.c_str(), (long long)h->m->N(), (long long)h->m->batch(), (long long)h->m->nvar(),
        (long long)h->m->ncon(), (long long)h->m->nnzj(), (long long)h->m->nnzh(), (long long)st.n[CTD_NEVAL_OBJ].load(),
        (long long)st.n[CTD_NEVAL_GRAD].load(), (long long)st.n[CTD_NEVAL_CONS].load(), (long long)st.n[CTD_NEVAL_JAC].load(),
        (long long)st.n[CTD_NEVAL_HESS].load(), (long long)st.n[CTD_NCALL_EVAL].load(), (long long)st.n[CTD_NCALL_HOST].load(),
        (long long)st.n[CTD_NCALL_NORMS].load(), (long long)h->m->launches(), 1e-9 * (double)st.host_ns.load());
    if (buf && buflen > 0) {
        const size_t k = std::min<size_t>((size_t)n, (size_t)buflen - 1);
        std::memcpy(buf, tmp, k);
        buf[k] = 0;
    }
    return (int64_t)n + 1;
}

int64_t ctd_kkt_nnz(ctd_handle h) {
    if (check(h)) return -1;
    int64_t n = -1;
    guarded([&] { n = h->m->kkt_nnz(); });
    return n;
}

int ctd_kkt_structure(ctd_handle h, int64_t* colptr, int64_t* rows) {
    if (int e = check(h)) return e;
    return guarded([&] { h->m->kkt_structure(colptr, rows); });
}

int ctd_kkt_assemble_dev(ctd_handle h, const double* hvals, const double* jvals, const double* dx, const double* dc,
                         double* kvals) {
    if (int e = check(h)) return e;
    return guarded([&] { h->m->kkt_assemble_dev(hvals, jvals, dx, dc, kvals); });
}

int ctd_nccl_unique_id(char id[CTD_NCCL_ID_BYTES]) {
    if (!id) return fail(-2, "ctd_nccl_unique_id: NULL argument");
    return guarded([&] { nccl_unique_id(id); });
}

int ctd_nccl_version(void) {
    int v = 0;
    const int e = guarded([&] { v = nccl_version(); });
    return e ? (e > 0 ? -e : e) : v;
}

int ctd_comm_init(ctd_handle h, const char id[CTD_NCCL_ID_BYTES], int32_t rank, int32_t world) {
    if (int e = check(h)) return e;
    return guarded([&] { h->m->comm_init(id, rank, world); });
}

int ctd_gather_norms_dev(ctd_handle h, const double* local, double* gathered, int32_t slot) {
    if (int e = check(h)) return e;
    return guarded([&] { h->m->gather_norms_dev(local, gathered, slot); });
}

int ctd_comm_wait(ctd_handle h, int32_t slot) {
    if (int e = check(h)) return e;
    return guarded([&] { h->m->comm_wait(slot); });
}

int ctd_peer_init(ctd_handle h) {
    if (int e = check(h)) return e;
    return guarded([&] { h->m->peer_init(); });
}

int ctd_eval_gather_dev(ctd_handle h, int what, const double* x, const double* y, double obj_weight, double* g,
                        double* c, double* jvals, double* hvals, int32_t slot) {
    if (int e = check(h)) return e;
    h->stats.count(what, false, true);
    return guarded([&] { h->m->eval_gather_dev(what, x, y, obj_weight, g, c, jvals, hvals, slot); });
}

int ctd_peer_wait(ctd_handle h, int32_t slot, int32_t overlapped, const double** gathered) {
    if (int e = check(h)) return e;
    return guarded([&] {
        const double* p = h->m->peer_wait(slot, overlapped != 0);
        if (gathered) *gathered = p;
    });
}

}  // extern "C"
