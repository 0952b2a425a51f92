/* ctdirect_b200.h -- C ABI of the B200-native NLP-evaluation engine for CTDirect.jl's discretised
 * optimal control problem (DOCP).
 *
 * The library (libctdirect_b200.so) replaces, for one discretised OCP, what CTDirect.jl obtains from
 * ADNLPModels on the CPU: the NLPModels callbacks that Ipopt / MadNLP call every iteration.  Each entry
 * point names the reference interface it stands for (paths relative to the CTDirect.jl tree; [3P] marks
 * the NLPModels API implemented by the un-vendored ADNLPModels package that CTDirect wires up in
 * src/collocation.jl:90-153).
 *
 * Conventions
 *  - return value: 0 = ok, < 0 = usage error, > 0 = CUDA error code (1000 + ncclResult_t for a failure of the
 *    norm gather); ctd_last_error() gives the message
 *    (thread-local).  Nothing is thrown across the boundary.  NaN/Inf in results are passed through.
 *  - all floating-point data are IEEE double; all indices are int64 and 1-BASED, in the order of
 *    Julia's CSC storage (column by column, rows ascending); the Hessian is the lower triangle.
 *  - buffers are owned by the caller and never retained after the call returns.
 *  - functions without suffix take HOST pointers (copied through device buffers owned by the handle,
 *    and synchronous); the `_dev` twins take DEVICE pointers, enqueue on the handle's stream and return
 *    without synchronising (ctd_sync waits).
 *  - batched handles (desc.batch = B > 1) evaluate B independent instances of the same DOCP that
 *    differ by their parameters (MPC scenarios); every array then holds B consecutive instances.
 *  - there is no CPU fallback: evaluation entry points fail with a CUDA error without a GPU.  The
 *    structural queries (dims, bounds, x0, structures) are host-only and work anywhere.
 */
#ifndef CTDIRECT_B200_H
#define CTDIRECT_B200_H

#include <stdint.h>

#if defined(__GNUC__)
#define CTD_API __attribute__((visibility("default")))
#else
#define CTD_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ctd_model_s* ctd_handle;

/* pattern_mode */
#define CTD_PATTERN_REFERENCE 0 /* bit-exact `:manual` patterns of the reference, quirks included */
#define CTD_PATTERN_CORRECTED 1 /* adds the entries the reference omits        */
#define CTD_PATTERN_OPTIMIZED 2 /* detected structural pattern (the reference's `:optimized` backend:
                                   SparseConnectivityTracer over the closures, collocation.jl:131-134);
                                   compiled for every (problem, scheme) pair                       */

/* output selection bits of ctd_eval / ctd_eval_dev */
#define CTD_CONS 1
#define CTD_JAC 2
#define CTD_HESS 4
#define CTD_GRAD 8
#define CTD_OBJ 16

typedef struct ctd_desc {
    const char* problem;      /* name of a compiled-in OCP ("goddard", "double_integrator_minenergy", ...)   */
    const char* scheme;       /* "trapeze" | "midpoint" | "gauss_legendre_{2,3}" |
                                 "gauss_legendre_{1,2,3}_constant_control"      (src/DOCP_data.jl:307-349)  */
    int32_t grid_size;        /* number of time steps N (Collocation option grid_size, collocation.jl:20-42) */
    const double* time_grid;  /* optional explicit grid (n_time_grid points, increasing) or NULL             */
    int32_t n_time_grid;
    const double* params;     /* optional problem parameters: n_params = npar (shared) or batch*npar         */
    int32_t n_params;
    int32_t pattern_mode;     /* CTD_PATTERN_*                                                               */
    int32_t batch;            /* number of independent instances (>= 1)                                      */
    int32_t device;           /* CUDA device ordinal                                                         */
    int32_t control_steps;    /* controls per time step: 0 or 1 = collocation (collocation.jl:65); > 1 = the
                                 DirectShooting discretizer (src/direct_shooting.jl:55-71: DOCP(ocp, grid_size,
                                 control_steps, :midpoint, nothing)).  scheme must be "midpoint"; the only pattern
                                 is the detected one (pattern_mode is ignored: DirectShooting hands no :manual
                                 pattern to ADNLPModels, direct_shooting.jl:101-137).  Compiled for 2, 3 (every
                                 problem with a control) and 10 (goddard, double_integrator_minenergy, vanderpol) */
} ctd_desc;

/* DOCP(ocp, grid_size, control_steps, scheme, time_grid) + __variables_bounds! + __constraints_bounds!
 * + the :manual pattern providers: src/DOCP_data.jl:293-365, src/collocation.jl:57-73,116-125.          */
CTD_API int ctd_create(const ctd_desc* desc, ctd_handle* out);
CTD_API int ctd_destroy(ctd_handle h);
CTD_API const char* ctd_last_error(void);
/* names of the compiled-in (problem, scheme, pattern_mode) combinations, one "problem scheme mode" per
 * line; returns the number of bytes needed (including the terminating NUL). */
CTD_API int64_t ctd_list_models(char* buf, int64_t buflen);

/* nlp.meta.{nvar, ncon, nnzj, nnzh}  [3P NLPModels meta]; dim_NLP_variables / dim_NLP_constraints of
 * src/DOCP_data.jl:276-277.  Per instance. */
CTD_API int ctd_dims(ctd_handle h, int64_t* nvar, int64_t* ncon, int64_t* nnzj, int64_t* nnzh);
/* out[0..15]: nx, nu, nv, path_cons, boundary_cons, N, step_variables_block, step constraint block,
 * batch, free_t0, free_tf, lagrange, mayer, max (1 => minimize=false, collocation.jl:145), npar, stage  */
CTD_API int ctd_info(ctd_handle h, int64_t out[16]);
/* nlp.meta.{lvar, uvar, lcon, ucon}: src/DOCP_variables.jl:21-63, src/DOCP_functions.jl:163-191 */
CTD_API int ctd_bounds(ctd_handle h, double* lvar, double* uvar, double* lcon, double* ucon);
/* nlp.meta.x0 for the default initial guess: __initial_guess, src/DOCP_variables.jl:122-145 */
CTD_API int ctd_x0(ctd_handle h, double* x0);
/* normalized time grid (N+1 values in [0,1]): DOCPtime, src/DOCP_data.jl:176-214 */
CTD_API int ctd_time_grid(ctd_handle h, double* s);
/* jac_structure!(nlp, rows, cols) [3P] for DOCP_Jacobian_pattern (src/ode/{trapeze,midpoint,irk,
 * irk_stagewise}.jl); nnzj entries, per instance */
CTD_API int ctd_jac_structure(ctd_handle h, int64_t* rows, int64_t* cols);
/* hess_structure!(nlp, rows, cols) [3P] for tril(DOCP_Hessian_pattern); nnzh entries */
CTD_API int ctd_hess_structure(ctd_handle h, int64_t* rows, int64_t* cols);

/* obj(nlp, x) [3P] -> __objective(x, docp), src/DOCP_functions.jl:23-54.  f has `batch` entries. */
CTD_API int ctd_obj(ctd_handle h, const double* x, double* f);
/* grad!(nlp, x, g) [3P] (ReverseDiffADGradient of __objective in the reference, collocation.jl:126-130) */
CTD_API int ctd_grad(ctd_handle h, const double* x, double* g);
/* cons!(nlp, x, c) [3P] -> __constraints!(c, x, docp), src/DOCP_functions.jl:80-115 */
CTD_API int ctd_cons(ctd_handle h, const double* x, double* c);
/* jac_coord!(nlp, x, vals) [3P] (SparseADJacobian in the reference, collocation.jl:116-120) */
CTD_API int ctd_jac_coord(ctd_handle h, const double* x, double* vals);
/* hess_coord!(nlp, x, y, vals; obj_weight) [3P]: Hessian of obj_weight*f(x) + y'c(x)
 * (SparseReverseADHessian in the reference, collocation.jl:121-125) */
CTD_API int ctd_hess_coord(ctd_handle h, const double* x, const double* y, double obj_weight, double* vals);
/* any subset of the above in ONE kernel launch; `what` is an OR of CTD_* bits out of
 * {CONS, JAC, HESS, OBJ, GRAD|OBJ, CONS|JAC|HESS, all five}; unused outputs may be NULL */
CTD_API int ctd_eval(ctd_handle h, int what, const double* x, const double* y, double obj_weight, double* f, double* g,
             double* c, double* jvals, double* hvals);

/* ctd_eval without the final wait: copies and kernel are only ENQUEUED on the handle's stream; the host
 * buffers (pinned memory, for the copies to be asynchronous) must stay untouched until ctd_sync(h).  Several
 * handles used round-robin overlap the H2D, kernel and D2H stages of consecutive evaluations. */
CTD_API int ctd_eval_async(ctd_handle h, int what, const double* x, const double* y, double obj_weight, double* f,
                   double* g, double* c, double* jvals, double* hvals);

/* device-pointer twins (asynchronous on the handle's stream) */
CTD_API int ctd_eval_dev(ctd_handle h, int what, const double* x, const double* y, double obj_weight, double* f,
                 double* g, double* c, double* jvals, double* hvals);
/* ctd_eval_dev plus the per-instance constraint violation max_k max(lcon_k - c_k, c_k - ucon_k, 0) reduced in
 * the same launch (`what` must include CTD_CONS; viol has `batch` entries) */
CTD_API int ctd_eval_norms_dev(ctd_handle h, int what, const double* x, const double* y, double obj_weight, double* f,
                       double* g, double* c, double* jvals, double* hvals, double* viol);
/* per-instance objective and constraint violation  max_k max(lcon_k - c_k, c_k - ucon_k, 0)  (the two
 * numbers a batched MPC driver gathers across GPUs); obj and viol have `batch` entries (device) */
CTD_API int ctd_norms_dev(ctd_handle h, const double* x, double* obj, double* viol);
CTD_API int ctd_norms(ctd_handle h, const double* x, double* obj, double* viol);
/* violation only, from constraint values c already on the device (e.g. produced by ctd_eval_dev) */
CTD_API int ctd_violation_dev(ctd_handle h, const double* c, double* viol);
/* Solution unpacking: getter(nlp_solution, docp; val) of src/ode/common.jl:7-104 with the normalisations
 * of build_OCP_solution (src/DOCP_data.jl:514-633).  Arrays are [node][component] (the memory of the
 * getter's matrices), one instance after the other; any member may be NULL.  y = NULL skips the duals.
 * The box multipliers (state_l/u, control_l/u, variable_l/u) use the primal layout: pass multipliers_L
 * or multipliers_U as x (with time = NULL). */
typedef struct ctd_solution {
    double* time;          /* (N+1)        time grid t0 + s (tf - t0), DOCP_data.jl:437-458          */
    double* state;         /* (N+1) x nx   val = :state                                               */
    double* control;       /* (N * control_steps + 1) x nu   val = :control (final control duplicated where aliased) */
    double* variable;      /* nv           val = :variable                                            */
    double* costate;       /* N x nx       val = :costate: multipliers of the state equations         */
    double* path_dual;     /* (N+1) x np   val = :mult_path_constraints divided by the step length    */
    double* boundary_dual; /* nb           val = :mult_boundary_constraints                           */
} ctd_solution;
CTD_API int ctd_unpack(ctd_handle h, const double* x, const double* y, const ctd_solution* out);
CTD_API int ctd_unpack_dev(ctd_handle h, const double* x, const double* y, const ctd_solution* out);
/* controls per time step of the handle (1 for collocation) and T_control of build_OCP_solution
 * (src/DOCP_data.jl:555-566): control_time[N * control_steps + 1] from the time[N+1] that ctd_unpack returned
 * (host arrays, one instance after the other). */
CTD_API int32_t ctd_control_steps(ctd_handle h);
CTD_API int ctd_control_time(ctd_handle h, const double* time, double* control_time);
/* Initial guess / warm start, the inverse of ctd_unpack: __initial_guess (src/DOCP_variables.jl:122-145,
 * irk_stagewise.jl:302-335) from SAMPLED data -- x = 0.1 everywhere, then variable, state[(N+1) x nx] at the
 * nodes and control[(N+1) x nu] through the setters of src/ode/common.jl:184-225 (final control only where
 * the scheme has one).  Stagewise schemes take stage_control[N x s x nu] (one control per stage) or, if it is
 * NULL, the node control for every stage of the step.  Any input may be NULL (its 0.1 default stays). */
CTD_API int ctd_pack(ctd_handle h, const double* state, const double* control, const double* variable,
             const double* stage_control, double* x);
CTD_API int ctd_pack_dev(ctd_handle h, const double* state, const double* control, const double* variable,
                 const double* stage_control, double* x);
/* replace the problem parameters (n = npar or batch*npar), e.g. new MPC scenarios */
CTD_API int ctd_set_params(ctd_handle h, const double* params, int32_t n);
CTD_API int ctd_sync(ctd_handle h);
/* cudaStream_t of the handle (for callers that enqueue their own work around the _dev calls) */
CTD_API void* ctd_stream(ctd_handle h);
/* Run the handle's work on a caller-owned cudaStream_t instead (NULL: back to the handle's own stream).  Work
 * already enqueued is finished first.  With this the evaluations take part in the caller's stream order and can
 * be captured into the caller's CUDA graphs (the _dev entry points enqueue kernels only: no allocation, no
 * synchronisation after the first call of each output combination).  The stream must outlive the handle or be
 * reset with NULL before it is destroyed. */
CTD_API int ctd_set_stream(ctd_handle h, void* stream);
/* number of kernel launches issued through this handle so far */
CTD_API int64_t ctd_launch_count(ctd_handle h);

/* Per-handle call counters: what NLPModels.Counters hold for the reference's models (neval_obj, neval_grad, neval_cons,
 * neval_jac, neval_hess; printed by the reference's benchmark, test/benchmark.jl:81-125).  Every entry point that
 * produces an output counts it -- a fused ctd_eval(CTD_CONS | CTD_JAC | CTD_HESS) adds one to each of the three.
 * ctd_stats_json writes the counters, the model's sizes, the kernel launches and the wall time spent inside the
 * blocking host-buffer calls as one JSON object; it returns the size needed (with the final NUL), like
 * ctd_list_models. */
enum {
    CTD_NEVAL_OBJ = 0, CTD_NEVAL_GRAD = 1, CTD_NEVAL_CONS = 2, CTD_NEVAL_JAC = 3, CTD_NEVAL_HESS = 4,
    CTD_NCALL_EVAL = 5,  /* evaluation calls of any kind */
    CTD_NCALL_HOST = 6,  /* of which with host buffers */
    CTD_NCALL_NORMS = 7, /* of which with the per-instance violation norm (batched) */
    CTD_N_COUNTERS = 8
};
CTD_API int ctd_counters(ctd_handle h, int64_t out[CTD_N_COUNTERS]);
CTD_API int ctd_reset_counters(ctd_handle h);
CTD_API int64_t ctd_stats_json(ctd_handle h, char* buf, int64_t buflen);

/* ---- KKT assembly.  Lower triangle, CSC, of
 *        K = [ H + diag(dx)       .      ]      (nvar + ncon columns)
 *            [ J              -diag(dc)  ]
 * the matrix an interior-point solver factorises (in the reference MadNLP assembles it from hess_coord! /
 * jac_coord! and hands it to cuDSS, test/test_gpu.jl:18).  Column j <= nvar: diagonal entry (always present),
 * strictly lower Hessian entries, Jacobian entries at rows nvar + r; column nvar + r: its diagonal entry.
 * ctd_kkt_structure: colptr (nvar+ncon+1) and rows (ctd_kkt_nnz), both 1-based.  ctd_kkt_assemble_dev moves the
 * value arrays of an evaluation (device pointers, [batch][nnz]) to KKT order in one launch; dx / dc (device,
 * [batch][nvar] / [batch][ncon]) may be NULL (zeros). */
CTD_API int64_t ctd_kkt_nnz(ctd_handle h);
CTD_API int ctd_kkt_structure(ctd_handle h, int64_t* colptr, int64_t* rows);
CTD_API int ctd_kkt_assemble_dev(ctd_handle h, const double* hvals, const double* jvals, const double* dx,
                                 const double* dc, double* kvals);

/* ---- multi-GPU norm gather (the reference has no batching, every scenario is its own DOCP,
 * src/DOCP_data.jl:293).  One process per GPU; each rank owns a batched handle for its shard of the
 * instances (all shards the same size).  The only exchange is the all-gather of the per-instance
 * (objective, violation) pairs over NCCL, issued by the library on a stream of its own.
 *   rank 0: ctd_nccl_unique_id(id); the caller hands the 128 bytes to every rank (MPI, a file, a TCP store...)
 *   every rank: ctd_comm_init(h, id, rank, world)
 *   per evaluation k, slot s = k % CTD_COMM_SLOTS:
 *     ctd_comm_wait(h, s)                      -- the gather that last used slot s has landed (frees local_s)
 *     ctd_eval_norms_dev(h, ..., f = local_s, ..., viol = local_s + batch)
 *     ctd_gather_norms_dev(h, local_s, gathered_s, s)
 *   before reading gathered_s on the handle's stream: ctd_comm_wait(h, s) (then ctd_sync for the host).
 * local_s: 2*batch doubles [objective | violation]; gathered_s: world*2*batch doubles, rank-major.
 * Waiting right after every gather serialises kernel and collective; waiting only before the slot is
 * reused overlaps the gather of evaluation k with the kernel of evaluation k+1. */
#define CTD_COMM_SLOTS 4
#define CTD_NCCL_ID_BYTES 128
CTD_API int ctd_nccl_unique_id(char id[CTD_NCCL_ID_BYTES]);
CTD_API int ctd_nccl_version(void);   /* e.g. 22809, or a negative error code when NCCL cannot be loaded */
CTD_API int ctd_comm_init(ctd_handle h, const char id[CTD_NCCL_ID_BYTES], int32_t rank, int32_t world);
CTD_API int ctd_gather_norms_dev(ctd_handle h, const double* local, double* gathered, int32_t slot);
CTD_API int ctd_comm_wait(ctd_handle h, int32_t slot);

/* ---- fused gather over peer memory (NVLink): the same exchange without a collective call.  After ctd_comm_init,
 * ctd_peer_init maps a gather buffer of every rank into every other rank (CUDA IPC handles exchanged over the
 * communicator; all ranks call it).  ctd_eval_gather_dev is ctd_eval_norms_dev whose reduction epilogue stores the
 * (objective, violation) of every instance straight into slot `slot` of EVERY rank's buffer and raises this rank's
 * arrival flag there when its shard is complete (CTD_CONS and CTD_OBJ must be among the outputs; light elements).
 * ctd_peer_wait enqueues the consumer -- a one-CTA kernel that polls the slot's records of all ranks until they
 * validate, copies them out and acknowledges the round to every rank -- and returns the slot's rows: world x 2 x batch
 * doubles, rank-major like ctd_gather_norms_dev.  overlapped = 0: on the handle's stream (what follows on the stream
 * sees the rows: the serial schedule); overlapped = 1: on the handle's communication stream; the work issued so
 * far is NOT awaited: the records are keyed by the launch, the consumer may poll from any time on, and the evaluation
 * stream keeps its evaluations back to back -- ctd_comm_wait(h, slot) joins it (the pipelined schedule).
 * A slot may be pushed again only after its ctd_peer_wait has been issued on every rank (a launch that pushes waits,
 * in its first CTA, for the acknowledgements of the slot's previous round; 4 slots).  A rank that never delivers is
 * reported by ctd_sync after 5 s.  Not capturable into CUDA graphs (the launch carries an epoch and a key). */
CTD_API int ctd_peer_init(ctd_handle h);
CTD_API int ctd_eval_gather_dev(ctd_handle h, int what, const double* x, const double* y, double obj_weight, double* g,
                                double* c, double* jvals, double* hvals, int32_t slot);
CTD_API int ctd_peer_wait(ctd_handle h, int32_t slot, int32_t overlapped, const double** gathered);

#ifdef __cplusplus
}
#endif
#endif /* CTDIRECT_B200_H */
