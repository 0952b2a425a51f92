/* The call sequence of ctdirect.jl_b200/julia/CTDirectB200.jl (b200_model_builder + the NLPModels callbacks), in C.
 *
 * Julia is not available in the build image, so the shim cannot be executed there; this program is compiled against
 * include/ctdirect_b200.h with a plain C compiler and performs exactly what the shim does for one (problem, scheme):
 *   ctd_create with the NORMALIZED time grid (docp.time.normalized_grid = collect(LinRange(0, 1, N+1)),
 *   src/DOCP_data.jl:176-214) -> ctd_dims -> ctd_info -> ctd_bounds -> ctd_x0 -> ctd_jac_structure / ctd_hess_structure
 *   -> ctd_obj, ctd_grad, ctd_cons, ctd_jac_coord, ctd_hess_coord (host pointers, as Ipopt calls them) -> ctd_eval
 *   (the fused call) -> ctd_unpack.
 * It writes the results of the separate callbacks AND of the fused call to a binary file that
 * tests/test_julia_shim_sequence.py compares with the oracle (the two come from different kernels: same values within
 * rounding, checked here only coarsely).
 *
 * With a 9th argument (control_steps > 1) it is the DirectShooting discretizer (docp.time.control_steps goes into
 * ctd_desc.control_steps).  At the end the library's own call counters (library_counters in the shim) must say what
 * was called.
 *
 *   julia_shim_sequence <problem> <scheme> <N> <pattern_mode> <x.bin> <y.bin> <out.bin> [control_steps]
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "ctdirect_b200.h"

#define CHECK(call)                                                                   \
    do {                                                                              \
        int e_ = (call);                                                              \
        if (e_ != 0) {                                                                \
            fprintf(stderr, "%s failed [%d]: %s\n", #call, e_, ctd_last_error());     \
            return 2;                                                                 \
        }                                                                             \
    } while (0)

static double* read_doubles(const char* path, int64_t n) {
    double* a = (double*)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
    FILE* f = fopen(path, "rb");
    if (!f || fread(a, sizeof(double), (size_t)n, f) != (size_t)n) {
        fprintf(stderr, "cannot read %lld doubles from %s\n", (long long)n, path);
        exit(3);
    }
    fclose(f);
    return a;
}

static int close_enough(const double* a, const double* b, int64_t n) {   /* coarse: 1e-9 of the largest magnitude */
    double m = 0.0, e = 0.0;
    for (int64_t k = 0; k < n; ++k) {
        if (fabs(a[k]) > m) m = fabs(a[k]);
        if (fabs(a[k] - b[k]) > e) e = fabs(a[k] - b[k]);
    }
    return e <= 1e-9 * (m > 1.0 ? m : 1.0);
}

int main(int argc, char** argv) {
    if (argc != 8 && argc != 9) {
        fprintf(stderr, "usage: %s problem scheme N pattern_mode x.bin y.bin out.bin\n", argv[0]);
        return 1;
    }
    const int N = atoi(argv[3]);
    /* docp.time.normalized_grid: collect(LinRange(0, 1, N + 1)) -- NOT docp.time.fixed_grid, which the reference
     * leaves uninitialised for free-time problems (src/DOCP_data.jl:203-205) */
    double* grid = (double*)malloc(sizeof(double) * (size_t)(N + 1));
    for (int j = 0; j <= N; ++j) grid[j] = (double)j / (double)N;
    ctd_desc d;
    memset(&d, 0, sizeof(d));
    d.problem = argv[1];
    d.scheme = argv[2];
    d.grid_size = N;
    d.time_grid = grid;
    d.n_time_grid = N + 1;
    d.pattern_mode = atoi(argv[4]);
    d.batch = 1;
    d.device = 0;
    d.control_steps = argc == 9 ? atoi(argv[8]) : 1;
    ctd_handle h = NULL;
    CHECK(ctd_create(&d, &h));
    int64_t nvar, ncon, nnzj, nnzh, info[16];
    CHECK(ctd_dims(h, &nvar, &ncon, &nnzj, &nnzh));
    CHECK(ctd_info(h, info));
    double *lvar = malloc(8 * nvar), *uvar = malloc(8 * nvar), *lcon = malloc(8 * (ncon + 1)), *ucon = malloc(8 * (ncon + 1));
    CHECK(ctd_bounds(h, lvar, uvar, lcon, ucon));
    double* x0 = malloc(8 * nvar);
    CHECK(ctd_x0(h, x0));
    int64_t *jr = malloc(8 * (nnzj + 1)), *jc = malloc(8 * (nnzj + 1)), *hr = malloc(8 * (nnzh + 1)), *hc = malloc(8 * (nnzh + 1));
    CHECK(ctd_jac_structure(h, jr, jc));
    CHECK(ctd_hess_structure(h, hr, hc));
    double* x = read_doubles(argv[5], nvar);
    double* y = read_doubles(argv[6], ncon);
    const double sigma = 0.7;
    /* the separate callbacks of NLPModels: obj, grad!, cons!, jac_coord!, hess_coord! */
    double f = 0.0, *g = malloc(8 * nvar), *c = malloc(8 * (ncon + 1)), *jv = malloc(8 * (nnzj + 1)), *hv = malloc(8 * (nnzh + 1));
    CHECK(ctd_obj(h, x, &f));
    CHECK(ctd_grad(h, x, g));
    CHECK(ctd_cons(h, x, c));
    CHECK(ctd_jac_coord(h, x, jv));
    CHECK(ctd_hess_coord(h, x, y, sigma, hv));
    /* eval_all!: one fused launch */
    double f2 = 0.0, *g2 = malloc(8 * nvar), *c2 = malloc(8 * (ncon + 1)), *jv2 = malloc(8 * (nnzj + 1)), *hv2 = malloc(8 * (nnzh + 1));
    CHECK(ctd_eval(h, CTD_CONS | CTD_JAC | CTD_HESS | CTD_GRAD | CTD_OBJ, x, y, sigma, &f2, g2, c2, jv2, hv2));
    if (!(close_enough(&f, &f2, 1) && close_enough(g, g2, nvar) && close_enough(c, c2, ncon) && close_enough(jv, jv2, nnzj) &&
          close_enough(hv, hv2, nnzh))) {
        fprintf(stderr, "the fused call and the separate callbacks disagree\n");
        return 4;
    }
    /* unpack_solution: time, state, control, variable, costate, path_dual, boundary_dual */
    const int64_t nx = info[0], nu = info[1], nv = info[2], np = info[3], nb = info[4];
    const int64_t ncs = ctd_control_steps(h);
    if (ncs != (d.control_steps > 1 ? d.control_steps : 1)) {
        fprintf(stderr, "ctd_control_steps says %lld\n", (long long)ncs);
        return 6;
    }
    double *tm = calloc((size_t)N + 1, 8), *st = calloc((size_t)(N + 1) * (nx ? nx : 1), 8), *ct = calloc((size_t)(N * ncs + 1) * (nu ? nu : 1), 8),
           *vr = calloc((size_t)(nv ? nv : 1), 8), *cs = calloc((size_t)N * (nx ? nx : 1), 8), *pd = calloc((size_t)(N + 1) * (np ? np : 1), 8),
           *bd = calloc((size_t)(nb ? nb : 1), 8);
    ctd_solution sol;
    memset(&sol, 0, sizeof(sol));
    sol.time = tm; sol.state = st; sol.control = nu ? ct : NULL; sol.variable = nv ? vr : NULL; sol.costate = cs;
    sol.path_dual = np ? pd : NULL; sol.boundary_dual = nb ? bd : NULL;
    CHECK(ctd_unpack(h, x, y, &sol));
    double* tc = calloc((size_t)(N * ncs + 1), 8);
    CHECK(ctd_control_time(h, tm, tc));
    if (tc[0] != tm[0] || tc[N * ncs] != tm[N] || tc[ncs] != tm[1]) {
        fprintf(stderr, "ctd_control_time does not interpolate the time grid\n");
        return 7;
    }
    /* the library's counters: obj, grad (= objective + gradient), cons, jac_coord, hess_coord, one fused call */
    int64_t cnt[CTD_N_COUNTERS];
    CHECK(ctd_counters(h, cnt));
    if (cnt[CTD_NEVAL_OBJ] != 3 || cnt[CTD_NEVAL_GRAD] != 2 || cnt[CTD_NEVAL_CONS] != 2 || cnt[CTD_NEVAL_JAC] != 2 ||
        cnt[CTD_NEVAL_HESS] != 2 || cnt[CTD_NCALL_EVAL] != 6 || cnt[CTD_NCALL_HOST] != 6) {
        fprintf(stderr, "ctd_counters: %lld %lld %lld %lld %lld calls %lld host %lld\n", (long long)cnt[0], (long long)cnt[1],
                (long long)cnt[2], (long long)cnt[3], (long long)cnt[4], (long long)cnt[5], (long long)cnt[6]);
        return 8;
    }
    FILE* o = fopen(argv[7], "wb");
    if (!o) return 5;
    int64_t hdr[8] = {nvar, ncon, nnzj, nnzh, info[13] /* max problem */, N, nx, nu};
    fwrite(hdr, 8, 8, o);
    fwrite(lvar, 8, nvar, o); fwrite(uvar, 8, nvar, o); fwrite(lcon, 8, ncon, o); fwrite(ucon, 8, ncon, o);
    fwrite(x0, 8, nvar, o);
    fwrite(jr, 8, nnzj, o); fwrite(jc, 8, nnzj, o); fwrite(hr, 8, nnzh, o); fwrite(hc, 8, nnzh, o);
    fwrite(&f, 8, 1, o); fwrite(g, 8, nvar, o); fwrite(c, 8, ncon, o); fwrite(jv, 8, nnzj, o); fwrite(hv, 8, nnzh, o);
    fwrite(tm, 8, (size_t)N + 1, o); fwrite(st, 8, (size_t)(N + 1) * nx, o);
    fwrite(&f2, 8, 1, o); fwrite(g2, 8, nvar, o); fwrite(c2, 8, ncon, o); fwrite(jv2, 8, nnzj, o); fwrite(hv2, 8, nnzh, o);
    fclose(o);
    CHECK(ctd_destroy(h));
    printf("ok %s %s N=%d nvar=%lld ncon=%lld nnzj=%lld nnzh=%lld\n", argv[1], argv[2], N, (long long)nvar, (long long)ncon,
           (long long)nnzj, (long long)nnzh);
    return 0;
}
