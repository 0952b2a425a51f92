// Fused DOCP evaluation kernels for sm_100a (FP64), templated on a generated element struct E
// (csrc/generated/*.cuh) and on the set F of requested outputs:
//   constraints c(x)            __constraints!      DOCP_functions.jl:80-115
//   Jacobian values             jac_coord!  on the :manual pattern, CSC order
//   Lagrangian-Hessian values   hess_coord! on the tril :manual pattern, CSC order
//   objective / gradient        __objective         DOCP_functions.jl:23-54
//
// Work decomposition:
//   * one thread evaluates one time step (all rows of the step and their analytic derivatives,
//     straight-line generated code);
//   * a WARP is an autonomous tile: lanes 1..31 own 31 consecutive *column blocks* i0..i0+30 of the
//     value arrays (the columns of the variables (x_i,u_i[,k_i]) of those steps) and lane 0
//     re-evaluates the halo step i0-1, whose derivatives w.r.t. x_{i0} land in block i0.  Warps only
//     synchronise with __syncwarp, so the load / compute / store phases of the warps resident on an SM
//     overlap each other;
//   * every value is written into the warp's shared-memory region at its final CSC position relative
//     to the tile, so the tile leaves the SM as a handful of contiguous ranges written with 16-byte
//     coalesced stores.  The region is used twice (phase A: constraints, Jacobian, gradient;
//     phase B: Hessian), which halves the shared memory per resident warp;
//   * the irregular ends (column blocks 1, N, N+1, boundary / final-path rows) are handled by CTA 0
//     through host-built position tables;
//   * (v,v) Hessian entries, d/dv of the cost and the Lagrange sum are reduced with warp shuffles into
//     per-CTA partials and summed in a fixed order by the last CTA to finish (deterministic, no
//     floating-point atomics on global memory).  Slots to which no STEP contributes (e.g. the (tf,tf)
//     entry of an autonomous free-final-time problem, a Mayer cost) are not reduced at all: the edge
//     CTA writes them;
//   * everything that does not depend on earlier work of the stream (tile geometry, normalized times,
//     parameters, edge tables, L2 prefetch of the tile's inputs) runs BEFORE griddepcontrol.wait, i.e.
//     while the previous kernel of the stream drains (programmatic dependent launch).
#pragma once
#include <cuda_runtime.h>

#include "ctd_elem_common.cuh"

namespace ctd {

constexpr int MAX_NV = 4;
constexpr int MAX_NVV = MAX_NV * (MAX_NV + 1) / 2;
constexpr int MAX_TASKS = 3;

enum : int { F_CONS = 1, F_JAC = 2, F_HESS = 4, F_GRAD = 8, F_OBJ = 16, F_VIOL = 32 };

struct EdgeTask {
    int step;      // 1-based step index
    int own;       // writes its own column block, its constraint rows and its v-column rows
    int map_off;   // offset of this task's map inside edge_map
    int next_rows; // E::PATH_SHIFT: writes the rows it holds of the NEXT step's block (the path rows of node step + 1)
};

struct EvalArgs {
    // per-instance arrays: instance b starts at pointer + b * stride
    const double* x;
    const double* y;
    double* c;
    double* jvals;
    double* hvals;
    double* grad;
    double* obj;
    const double* params;
    long long x_stride, y_stride, c_stride, j_stride, h_stride, g_stride;
    double sigma;
    const double* sgrid;   // normalized time grid (N+1 values in [0,1])
    int N;
    int tiles;             // warp tiles per instance
    int ctas;              // CTAs per instance: CTA 0 is the edge CTA, CTAs 1.. hold WPC warp tiles each
    long long nvar, ncon;
    long long jv_base[MAX_NV];    // position of the first value of column v_k in jvals
    long long hvv_pos[MAX_NVV];   // positions of the (v_a, v_b) entries in hvals (packed lower triangle)
    double* viol;                 // [batch] max(lcon - c, c - ucon, 0) per instance, or nullptr
    double* partials;             // [batch][ctas][NREDX]
    unsigned int* counters;       // [batch]
    // fused norm gather over peer memory (multi-GPU batches, ctd_eval_gather_dev; peer_base == nullptr otherwise): as
    // soon as an instance is reduced, its (objective, violation) go into the gather buffer of EVERY rank as two
    // self-validating 16-byte records {bits(v), bits(v) ^ key} (NVLink peer stores, posted: no fence, no flag, no
    // counter).  `key` changes with every launch, so the consumer (ctd_peer_wait_kernel) tells a record of this round
    // from what the slot held before, and a torn read never validates.
    double* const* peer_base;     // [peer_world] base of every rank's gather buffer (peer-mapped device pointers)
    long long peer_rec;           // offset (doubles) of this rank's records [2][batch] for the slot in use
    unsigned long long peer_key, peer_epoch;
    int peer_world;
    // slot recycling: a rank acknowledges a round when its consumer has copied the slot out (ctd_peer_wait_kernel);
    // the first CTA of the launch waits for every rank's acknowledgement of the slot's previous round in the LOCAL ack
    // table (normally long there) and opens the gate for the pushers of this launch
    const unsigned long long* peer_ack_local;
    unsigned long long* peer_gate;
    // edge tables (identical for every instance)
    const int* edge_map;
    const long long* edge_pos;    // destination position of every edge-buffer entry
    int ne_j, ne_h, ne_g;         // edge-buffer sections: J values, H values, gradient entries
    int ntask;
    EdgeTask task[MAX_TASKS];
    int fp_map_off, bc_map_off, nmap;
    int ip_map_off;   // E::PATH_SHIFT: map of the initial-path job (path rows of node 1), laid out like the final-path map
    long long* trace;             // debug builds (-DCTD_TRACE): per-warp timestamps, else nullptr
};

// ---------------------------------------------------------------------------------------------
template <class E, int F>
struct Cfg {
    static constexpr bool WC = (F & F_CONS) != 0, WJ = (F & F_JAC) != 0, WH = (F & F_HESS) != 0,
                          WG = (F & F_GRAD) != 0, WO = (F & F_OBJ) != 0, WV = (F & F_VIOL) != 0;
    static_assert(!WV || WC, "the violation norm is a by-product of the constraint rows");
    static constexpr int NV = E::NV;
    static constexpr int NVV = NV * (NV + 1) / 2;
    // grid-wide sums exist only for the quantities a STEP contributes to: (v,v) Hessian entries (none for
    // autonomous dynamics with the free times entering through h only), d cost / d v and the cost itself
    // (none for a Mayer cost).  What only the boundary / final-path elements touch is written by the edge
    // CTA through its value buffer (Model::build_edge_tables makes the same distinction).
    static constexpr bool RH = WH && E::NH_VV_EMIT > 0, RG = WG && E::NG_V > 0, RO = WO && E::LAGRANGE;
    static constexpr int R_H = 0, R_G = R_H + (RH ? NVV : 0), R_O = R_G + (RG ? NV : 0);
    static constexpr int NRED = R_O + (RO ? 1 : 0);
    // one more slot, combined with max instead of +: the constraint violation of the instance
    static constexpr int R_V = NRED, NREDX = NRED + (WV ? 1 : 0);
    static constexpr int NJV = NV * E::JV_STRIDE;
    // phase A stages constraints, Jacobian and gradient; phase B the Hessian (same shared memory)
    static constexpr bool PHASE_A = WC || WJ || WG || WO, PHASE_B = WH;
    // GZ: no step contributes to the objective gradient (Mayer cost only, e.g. Goddard, robot arm): there is
    // nothing to stage, the tiles store the zeros of their column blocks directly; SG: gradient staged
    static constexpr bool GZ = WG && E::NG_OWN == 0 && E::NG_NEXT == 0 && E::NG_V == 0;
    static constexpr bool SG = WG && !GZ;
    // Row stride of the staged Hessian block.  Every lane writes its own row, so a stride of H_BLOCK doubles puts the 16
    // lanes of a half-warp on (H_BLOCK mod 16) bank pairs: 104 doubles (3-stage IRK of a 3-state problem) = 8 mod 16, an
    // 8-way conflict on every one of the 104 stores (ncu: 1.4 M shared-memory bank conflicts per launch; phase B of the
    // parts kernel spends more time replaying stores than computing).  An odd stride is conflict-free, an even stride
    // 2 * odd costs two passes: even block sizes whose half is even are padded to the next 2 * odd, and the rows then
    // leave as one bulk copy each.
    static constexpr int pad_stride(int n) { return (n % 2 == 1 || (n / 2) % 2 == 1) ? n : n + 2; }
    static constexpr int HS = pad_stride(E::H_BLOCK);
    static constexpr bool H_PADDED = HS != E::H_BLOCK;
    // the same for the Jacobian block (stagewise 3-stage IRK of a 3-state problem: 192 doubles = 0 mod 16, a 16-way conflict)
    static constexpr int JS = pad_stride(E::J_BLOCK);
    static constexpr bool J_PADDED = JS != E::J_BLOCK;
    static constexpr int PER_STEP_A = (WC ? E::CB : 0) + (WJ ? JS + NJV : 0) + (SG ? E::BLK : 0);
    static constexpr int PER_STEP_B = WH ? HS : 0;
    static constexpr int PER_STEP = PER_STEP_A > PER_STEP_B ? PER_STEP_A : PER_STEP_B;
    // WPC = the LARGEST number of warps per CTA a launch may use (the host picks blockDim per launch,
    // see Model::eval_dev): 8 while a CTA's staged tiles stay below 72 KB, fewer for wide elements
    // (e.g. 3-stage IRK of a 6-state problem)
    static constexpr int SMEM_MAX = 200 * 1024;
#ifndef CTD_WPC
#define CTD_WPC 8
#endif
    static constexpr int WARP_BYTES = PER_STEP * 8 * 32;
#ifndef CTD_WPC_SMEM_KB
#define CTD_WPC_SMEM_KB 72
#endif
    static constexpr int WPC = (WARP_BYTES * CTD_WPC <= CTD_WPC_SMEM_KB * 1024) ? CTD_WPC
                               : (WARP_BYTES * 4 <= 100 * 1024) ? 4 : (WARP_BYTES * 2 <= 100 * 1024) ? 2 : 1;
    static constexpr int T = WPC * 32;
    // light elements (midpoint / trapeze / Euler of small systems): registers are capped at 80 so that 24
    // warps (744 owned steps) are resident per SM and a grid of 1e5 steps is ONE wave on 148 SMs; wide
    // elements keep the full register file
#ifndef CTD_MAXNREG
#define CTD_MAXNREG 80
#endif
    // (few staged values per step is not enough: the robot arm with the detected pattern stages 30 doubles
    //  per step but its derivative code spills 168 bytes under an 80-register cap, the 2-stage IRK of Goddard
    //  with the detected pattern 256 bytes)
#ifndef CTD_LIGHT_MAXDIM
#define CTD_LIGHT_MAXDIM 6
#endif
#ifndef CTD_LIGHT_MAXSTAGE
#define CTD_LIGHT_MAXSTAGE 1
#endif
    static constexpr bool LIGHT = PER_STEP <= 64 && (E::NX + E::NU) <= CTD_LIGHT_MAXDIM && E::STAGE <= CTD_LIGHT_MAXSTAGE;
    static constexpr int FIT = PER_STEP > 0 ? (SMEM_MAX / WPC - 512) / (PER_STEP * 8) - 1 : 31;
    static constexpr int OWN = (31 < FIT) ? 31 : FIT;  // owned steps per warp tile
    static_assert(OWN >= 1, "element too wide for one SM's shared memory");
    // sub-buffers of a warp's region (each with 2 doubles of slack so that the staged range can be
    // given the same 16-byte phase as its destination in global memory)
    static constexpr int even(int n) { return (n + 3) & ~1; }
    static constexpr int S_C = 0;
    // (OWN + 1 blocks each: the last one is a dump block, written by the halo lane's "own" values and
    // the last lane's "next" values, so that the sinks store without predicates or branches)
    static constexpr int S_J = S_C + (WC ? even((OWN + 1) * E::CB) : 0);
    static constexpr int S_JV = S_J + (WJ ? even((OWN + 1) * JS) : 0);
    static constexpr int JV_ONE = even((OWN + 1) * E::JV_STRIDE);
    static constexpr int S_G = S_JV + (WJ ? NV * JV_ONE : 0);
    static constexpr int A_END = S_G + (SG ? even((OWN + 1) * E::BLK) : 0);
    static constexpr int S_H = 0;
    static constexpr int B_END = WH ? even((OWN + 1) * HS) : 0;
    static constexpr int WARP_DOUBLES = ((A_END > B_END ? A_END : B_END) + 1) & ~1;
    static constexpr int S_RED = WPC * WARP_DOUBLES;
    static constexpr int S_END = S_RED + WPC * (NREDX > 0 ? NREDX : 1) + 2;
    static constexpr size_t TILE_SMEM = (size_t)S_END * sizeof(double);
    // edge map layout of one step task
    static constexpr int M_JO = 0, M_JN = M_JO + E::NJ_OWN, M_JV = M_JN + E::NJ_NEXT, M_HO = M_JV + E::NJ_V,
                         M_HN = M_HO + E::NH_OWN, M_GO = M_HN + E::NH_NEXT, M_GN = M_GO + E::NG_OWN,
                         M_STEP = M_GN + E::NG_NEXT;
    static constexpr int M_FP_J = 0, M_FP_H = M_FP_J + E::NFP_J, M_FP = M_FP_H + E::NFP_H;
    static constexpr int M_BC_J = 0, M_BC_H = M_BC_J + E::NBC_J, M_BC_G = M_BC_H + E::NBC_H, M_BC = M_BC_G + E::NBC_G;
};

// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double grid_point(const EvalArgs& a, long long j /*0-based node*/) {
    // DOCP_data.jl:176-214: the normalized grid (collect(LinRange(0,1,N+1)) by default) is built once on
    // the host; it is shared by all instances and evaluations, so it stays L2-resident
    return __ldg(a.sgrid + j);
}

// t_i = t0 + s_i * (tf - t0) for the two nodes of a step (normalized times si, sp), the step length and their
// forward-mode seeds w.r.t. the free times in v  (get_time_grid: DOCP_data.jl:437-458)
template <class E>
__device__ __forceinline__ void time_args_s(const double si, const double sp, const double* v, const double* p,
                                            TimeArgs<E::NV>& T) {
    const double t0 = E::T0_KIND == 0 ? E::T0_CONST : (E::T0_KIND == 1 ? v[E::T0_IDX > 0 ? E::T0_IDX : 0] : p[E::T0_IDX > 0 ? E::T0_IDX : 0]);
    const double tf = E::TF_KIND == 0 ? E::TF_CONST : (E::TF_KIND == 1 ? v[E::TF_IDX > 0 ? E::TF_IDX : 0] : p[E::TF_IDX > 0 ? E::TF_IDX : 0]);
    T.ti = CTD_ADD(t0, CTD_MUL(si, CTD_SUB(tf, t0)));
    T.tp = CTD_ADD(t0, CTD_MUL(sp, CTD_SUB(tf, t0)));
    T.hh = CTD_SUB(T.tp, T.ti);
#pragma unroll
    for (int k = 0; k < E::NV; ++k) {
        double di = 0.0, dp = 0.0;
        if (E::TF_KIND == 1 && k == E::TF_IDX) { di += si; dp += sp; }
        if (E::T0_KIND == 1 && k == E::T0_IDX) { di += 1.0 - si; dp += 1.0 - sp; }
        T.dti[k] = di;
        T.dtp[k] = dp;
        T.dh[k] = dp - di;
    }
}

template <class E>
__device__ __forceinline__ void time_args(const EvalArgs& a, const double* v, const double* p, long long i,
                                          TimeArgs<E::NV>& T) {
    const double si = grid_point(a, i - 1);
    const double sp = (i <= a.N) ? grid_point(a, i) : si;
    time_args_s<E>(si, sp, v, p, T);
}

// ---- TMA bulk store (cp.async.bulk, shared::cta -> global): one elected lane hands a staged range to
// the copy engine, the warp's issue slots stay free for the next phase.
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ unsigned int smem_u32(const void* p) { return (unsigned int)__cvta_generic_to_shared(p); }
// One lane of a converged warp, chosen by the hardware: ptxas then knows that the block it guards runs in a
// single thread and moves its operands to uniform registers without a per-value loop (a `lane == 0` test
// costs ~25 instructions per bulk copy).
__device__ __forceinline__ bool elect_one() {
    unsigned int pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ int phase_of(const double* g) {
    return (int)(((unsigned int)reinterpret_cast<unsigned long long>(g) >> 3) & 1u);
}
// A staged range of n doubles (n = 0: nothing -- a v column without per-step entries, an empty Hessian block of a
// detected pattern), destination g (8-byte aligned), staged copy s with the SAME 16-byte phase
// as g (s_u32 = its shared-memory address).  The 16-byte-aligned body goes through the copy engine (issued by
// the elected lane), the at most one leading and one trailing double are stored directly (by lane 1).
// The results stream out once and are not read again by the kernels: they are written with the L2 evict-first
// policy, so that what every launch re-reads (its own code, the time grid, parameters, edge tables) is not pushed out
// of the 126 MB L2 by the results of earlier evaluations.
#ifndef CTD_EVICT_FIRST
#define CTD_EVICT_FIRST 1
#endif
__device__ __forceinline__ unsigned long long evict_first_policy() {
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void bulk_body(double* __restrict__ g, unsigned int s_u32, int n) {
    if (n <= 0) return;
    const int ph = phase_of(g);
    const int body = (n - ph) >> 1;
    if (body > 0) {
#if CTD_EVICT_FIRST
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;"
                     :: "l"(g + ph), "r"(s_u32 + 8u * (unsigned int)ph), "r"(body << 4), "l"(evict_first_policy()) : "memory");
#else
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                     :: "l"(g + ph), "r"(s_u32 + 8u * (unsigned int)ph), "r"(body << 4) : "memory");
#endif
    }
}
__device__ __forceinline__ void ragged_ends(double* __restrict__ g, const double* __restrict__ s, int n) {
    if (n <= 0) return;
    const int ph = phase_of(g);
    if (ph) g[0] = s[0];
    if ((n - ph) & 1) g[n - 1] = s[n - 1];
}

#ifdef CTD_TRACE
__device__ __forceinline__ long long gtime() {
    long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
#define CTD_STAMP(k) do { if (a.trace && lane == 0) a.trace[((long long)cta * (blockDim.x >> 5) + warp) * 8 + (k)] = gtime(); } while (0)
// slot 5 (the "published" stamp) is replaced by the SM the warp ran on
#define CTD_STAMP_SMID() do { if (a.trace && lane == 0) { unsigned int sm_; asm volatile("mov.u32 %0, %%smid;" : "=r"(sm_)); \
    a.trace[((long long)cta * (blockDim.x >> 5) + warp) * 8 + 5] = (long long)sm_; } } while (0)
#else
#define CTD_STAMP_SMID() do { } while (0)
#define CTD_STAMP(k) do { } while (0)
#endif

// NaN-propagating max (a NaN constraint value must surface as a NaN violation)
__device__ __forceinline__ double nanmax(double a, double b) { return (b > a || b != b) ? b : a; }
__device__ __forceinline__ double row_violation(double lb, double ub, double v) { return nanmax(lb - v, v - ub); }

// ---------------------------------------------------------------------------------------------
// Sink of a tile thread for one phase (PC/PJ/PH/PG select what this phase keeps): derivative values go
// to their CSC slot inside the warp's staged tile.
template <class E, int F, bool PC, bool PJ, bool PH, bool PG>
struct TileSink {
    using C = Cfg<E, F>;
    double* c_;       // own constraint rows
    double* cn_;      // rows of the NEXT step's block (E::PATH_SHIFT: the path rows of node i+1, see gen/elements.py)
    double* jvn_[E::NV > 0 ? E::NV : 1];  // v columns, the next step's rows (E::PATH_SHIFT)
    double* jo_;      // own column block (Jacobian)
    double* jn_;      // next column block (Jacobian)
    double* jv_[E::NV > 0 ? E::NV : 1];  // v columns, this step's rows
    double* ho_;      // own column block (Hessian)
    double* go_;      // own gradient block
    bool owner, giver;
    double hn_[E::NH_NEXT > 0 ? E::NH_NEXT : 1];
    double gn_[E::NG_NEXT > 0 ? E::NG_NEXT : 1];
    double* red;

    template <int R> __device__ __forceinline__ void c(double v) {
        if (PC) {
            if constexpr (E::PATH_SHIFT && R >= E::EQS) {   // a row of the next step's block: the giver stores it
                cn_[R] = v;
                if constexpr (C::WV) { if (giver) red[C::R_V] = nanmax(red[C::R_V], row_violation(E::CON_LB(R), E::CON_UB(R), v)); }
            } else {
                c_[R] = v;
                if constexpr (C::WV) { if (owner) red[C::R_V] = nanmax(red[C::R_V], row_violation(E::CON_LB(R), E::CON_UB(R), v)); }
            }
        }
    }
    template <int N> __device__ __forceinline__ void jo(double v) { if (PJ) jo_[E::J_OWN(N, 2)] = v; }
    template <int N> __device__ __forceinline__ void jn(double v) { if (PJ) jn_[E::J_NEXT(N, 2)] = v; }
    template <int N> __device__ __forceinline__ void jv(double v) {
        if (PJ) {
            // slots >= JV_STRIDE lie in the next step's chunk of the column (shifted rows)
            if constexpr (E::PATH_SHIFT && E::J_V(N, 2) >= E::JV_STRIDE) jvn_[E::J_V(N, 1)][E::J_V(N, 2) - E::JV_STRIDE] = v;
            else jv_[E::J_V(N, 1)][E::J_V(N, 2)] = v;
        }
    }
    template <int N> __device__ __forceinline__ void ho(double v) { if (PH) ho_[E::H_OWN(N, 2)] = v; }
    template <int N> __device__ __forceinline__ void hn(double v) { if (PH) hn_[N] = v; }
    template <int K> __device__ __forceinline__ void hv(double v) { if constexpr (PH && C::RH) red[C::R_H + K] += owner ? v : 0.0; }
    template <int N> __device__ __forceinline__ void go(double v) { if (PG) go_[E::G_OWN(N, 2)] = v; }
    template <int N> __device__ __forceinline__ void gn(double v) { if (PG) gn_[N] = v; }
    template <int N> __device__ __forceinline__ void gv(double v) { if constexpr (PG && C::RG) red[C::R_G + E::G_V(N, 2)] += owner ? v : 0.0; }
};

// Sink of an edge step job: like a tile, every value is staged at its compile-time slot of the task's
// staging block (one STS each, no map lookups or branches in the derivative code); the whole CTA then
// places the staged values into the edge buffer through the host-built map (edge_cta below).
// map codes: >= 0 edge-buffer index, -1 not in the pattern (dropped), <= -2 reduction slot -(m+2).
template <class E, int F>
struct EdgeSink {
    using C = Cfg<E, F>;
    double* st;       // staging block of this task: C::M_STEP doubles, laid out like the task's map
    bool sums;        // this task contributes to the global sums (it is not a halo re-evaluation)
    double* crow;     // destination of the constraint rows (global) or nullptr
    double* crow_next;   // E::PATH_SHIFT: destination of the rows of the next step's block, or nullptr
    double* red;
    template <int R> __device__ __forceinline__ void c(double v) {
        double* dst = (E::PATH_SHIFT && R >= E::EQS) ? crow_next : crow;
        if (dst) {
            dst[R] = v;
            if constexpr (C::WV) red[C::R_V] = nanmax(red[C::R_V], row_violation(E::CON_LB(R), E::CON_UB(R), v));
        }
    }
    template <int N> __device__ __forceinline__ void jo(double v) { st[C::M_JO + N] = v; }
    template <int N> __device__ __forceinline__ void jn(double v) { st[C::M_JN + N] = v; }
    template <int N> __device__ __forceinline__ void jv(double v) { st[C::M_JV + N] = v; }
    template <int N> __device__ __forceinline__ void ho(double v) { st[C::M_HO + N] = v; }
    template <int N> __device__ __forceinline__ void hn(double v) { st[C::M_HN + N] = v; }
    template <int K> __device__ __forceinline__ void hv(double v) { if constexpr (C::RH) { if (sums) red[C::R_H + K] += v; } }
    template <int N> __device__ __forceinline__ void go(double v) { st[C::M_GO + N] = v; }
    template <int N> __device__ __forceinline__ void gn(double v) { st[C::M_GN + N] = v; }
    template <int N> __device__ __forceinline__ void gv(double v) { if constexpr (C::RG) { if (sums) red[C::R_G + E::G_V(N, 2)] += v; } }
};

// final path / boundary elements (edge CTA only): c<R>, j<N>, h<N>, g<N>
template <class E, int F, int MJ, int MH, int MG, bool BOUNDARY>
struct EdgeElemSink {
    using C = Cfg<E, F>;
    double* buf;
    const int* map;
    double* crow;
    double* red;
    template <int R> __device__ __forceinline__ void c(double v) {
        if (crow) {
            crow[R] = v;
            if constexpr (C::WV) {   // rows of the boundary block, or path-constraint rows at the final time
                const double lb = BOUNDARY ? E::BC_LB(R) : E::CON_LB(E::EQS + R);
                const double ub = BOUNDARY ? E::BC_UB(R) : E::CON_UB(E::EQS + R);
                red[C::R_V] = nanmax(red[C::R_V], row_violation(lb, ub, v));
            }
        }
    }
    template <int N> __device__ __forceinline__ void j(double v) {
        const int m = map[MJ + N];
        if (m >= 0) buf[m] = v;
    }
    template <int N> __device__ __forceinline__ void h(double v) {
        const int m = map[MH + N];
        if (m >= 0) buf[m] = v;
        else if constexpr (C::RH) { if (m <= -2) red[C::R_H + (-m - 2)] += v; }
    }
    template <int N> __device__ __forceinline__ void g(double v) {
        const int m = map[MG + N];
        if (m >= 0) buf[m] = v;
        else if constexpr (C::RG) { if (m <= -2) red[C::R_G + (-m - 2)] += v; }
    }
};

// ---------------------------------------------------------------------------------------------
template <class E>
__device__ __forceinline__ void load_step(const double* __restrict__ x, long long nvar, long long i,
                                          double* __restrict__ z) {
    const double* xo = x + (i - 1) * E::BLK;
#pragma unroll
    for (int k = 0; k < E::NOWN + E::NNEXT; ++k) z[k] = __ldg(xo + k);
#pragma unroll
    for (int k = 0; k < E::NV; ++k) z[E::NOWN + E::NNEXT + k] = __ldg(x + nvar - E::NV + k);
}

// Programmatic dependent launch: a launch made with the programmatic-stream-serialization attribute may
// be scheduled while the previous kernel of the stream is still draining; `griddepcontrol.wait` blocks
// until that kernel has completed and its memory is visible.  Everything a thread does BEFORE it must not
// depend on (or disturb) earlier work of the stream: tile geometry, the normalized time grid, parameters
// and edge tables (handle-owned, only ever written by stream-ordered copies), shared-memory setup, and L2
// prefetches of the inputs (a prefetch returns no data; the L2 is the coherence point).
#ifndef CTD_PDL
#define CTD_PDL 1
#endif

__device__ __forceinline__ void pdl_entry() {
#if CTD_PDL
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" :: "l"(p)); }

// Edge CTA: the elements that touch column blocks 1, N, N+1 and the boundary / final-path rows.  The
// work is cut into jobs -- (step elements, phase A), (step elements, phase B), final-path element,
// boundary+Mayer element -- one warp each; the step elements of one phase run in the lanes of one warp.  Jobs whose values can land on the same entry
// write to different layers of the edge buffer, which are summed in a fixed order when the buffer is
// scattered (deterministic whatever the warp interleaving, and no atomics).
// layers: step tasks 0/2, step task 1, final path, boundary (+ the initial-path job of E::PATH_SHIFT elements)
constexpr int edge_layers(bool path_shift) { return path_shift ? 5 : 4; }

template <class E, int F, bool PA, bool PB>
__device__ __forceinline__ void edge_step_job(const EvalArgs& a, const EdgeTask& tk, const double* __restrict__ x,
                                              const double* __restrict__ y, const double* __restrict__ p, double* c,
                                              double* stage, const int* smap, double* red) {
    using C = Cfg<E, F>;
    constexpr bool JC = PA && C::WC, JJ = PA && C::WJ, JH = PB && C::WH, JG = PA && C::WG, JO = PA && C::RO;
    double z[E::NZ], lam[E::CB];
    load_step<E>(x, a.nvar, tk.step, z);
#pragma unroll
    for (int r = 0; r < E::CB; ++r) {
        if (E::PATH_SHIFT && r >= E::EQS)   // multipliers of the next node's path rows; those of node N+1 belong to the final-path element
            lam[r] = (JH && y && tk.step < a.N) ? __ldg(y + (long long)tk.step * E::CB + r) : 0.0;
        else
            lam[r] = (JH && y) ? __ldg(y + (long long)(tk.step - 1) * E::CB + r) : 0.0;
    }
    TimeArgs<E::NV> Tm;
    time_args<E>(a, z + E::NOWN + E::NNEXT, p, tk.step, Tm);
    EdgeSink<E, F> s;
    s.st = stage;
    s.sums = smap[tk.map_off + C::M_STEP] != 0;
    s.crow = (JC && tk.own) ? c + (long long)(tk.step - 1) * E::CB : nullptr;
    s.crow_next = (E::PATH_SHIFT && JC && tk.next_rows) ? c + (long long)tk.step * E::CB : nullptr;
    s.red = red;
    if constexpr (JC || JJ || JH) E::template step<JC, JJ, JH>(z, Tm, p, lam, a.sigma, s);
    if constexpr (JG) E::grad_step(z, Tm, p, s);
    if constexpr (JO) { if (tk.own) red[C::R_O] += E::cost_step(z, Tm, p); }
}

// shared-memory layout of the edge CTA (dynamic shared memory, sized by Model::eval_dev):
//   EDGE_LAYERS x ne doubles (value layers) | nmap ints (position maps) | ntask x M_STEP doubles (staging
//   blocks of the step tasks) | ne int64 (destination positions)
template <class E, int F>
__device__ void edge_cta(const EvalArgs& a, const double* __restrict__ x, const double* __restrict__ y,
                         const double* __restrict__ p, double* c, double* jvals, double* hvals, double* grad,
                         double* obj, double* buf, double* red, const int NW) {
    using C = Cfg<E, F>;
    const int tid = threadIdx.x;
    const int T = NW * 32;
    const int ne = a.ne_j + a.ne_h + a.ne_g;
    constexpr int EDGE_LAYERS = edge_layers(E::PATH_SHIFT);
    int* smap = reinterpret_cast<int*>(buf + EDGE_LAYERS * ne);
    double* stage = buf + EDGE_LAYERS * ne + (a.nmap + 1) / 2;
    long long* spos = reinterpret_cast<long long*>(stage + a.ntask * C::M_STEP);
    // ---- before the grid dependency: the (immutable) position tables are staged in shared memory -- the
    //      sinks look up one map entry per value and those lookups must not be dependent global loads --
    //      and the value layers are cleared
    for (int k = tid; k < a.nmap; k += T) smap[k] = __ldg(a.edge_map + k);
    for (int k = tid; k < ne; k += T) spos[k] = __ldg(a.edge_pos + k);
    for (int k = tid; k < EDGE_LAYERS * ne; k += T) buf[k] = 0.0;
    pdl_entry();
    if (a.peer_base && blockIdx.y == 0 && tid < 32) {   // slot recycling of the fused gather (see EvalArgs)
        if (tid < a.peer_world) {
            for (int spin = 0; spin < (1 << 23); ++spin) {   // (bounded: ~2 s; a dead rank must not hang this one)
                unsigned long long ak;
                asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(ak) : "l"(a.peer_ack_local + tid) : "memory");
                if (ak + 1 >= a.peer_epoch) break;
                __nanosleep(200);
            }
        }
        __syncwarp();
        if (tid == 0) asm volatile("st.release.gpu.global.u64 [%0], %1;" :: "l"(a.peer_gate), "l"(a.peer_epoch) : "memory");
    }
    __syncthreads();
    {
        // warp roles: the step elements of all edge tasks run side by side in the lanes of one warp
        // (phase A on one warp, phase B on another when both are requested); the final-path and the
        // boundary element take lane 0 of the next warps (round robin when the CTA has fewer warps).
        constexpr bool TWO = C::PHASE_A && C::PHASE_B;
        const int W_A = 0, W_B = TWO ? (1 % NW) : 0, W_FP = (TWO ? 2 : 1) % NW, W_BC = (TWO ? 3 : 2) % NW;
        const int W_IP = (TWO ? 4 : 3) % NW;
        const int lane_ = tid & 31, warp_ = tid >> 5;
        if (lane_ < a.ntask) {
            const EdgeTask tk = a.task[lane_];
            double* st = stage + lane_ * C::M_STEP;
            if constexpr (TWO) {
                if (warp_ == W_A) edge_step_job<E, F, true, false>(a, tk, x, y, p, c, st, smap, red);
                if (warp_ == W_B) edge_step_job<E, F, false, true>(a, tk, x, y, p, c, st, smap, red);
            } else {
                if (warp_ == W_A) edge_step_job<E, F, C::PHASE_A, C::PHASE_B>(a, tk, x, y, p, c, st, smap, red);
            }
        }
        if (lane_ == 0) {
            if (warp_ == W_FP) {
                // final path constraints g(t_f, x_f, u_f, v)  (DOCP_functions.jl:100)
                if constexpr (E::NP > 0 && (C::WC || C::WJ || C::WH)) {
                    double z[E::FP_NLOC], lam[E::NP];
                    const long long N = a.N;
                    if (E::FINAL_CONTROL) {
#pragma unroll
                        for (int k = 0; k < E::NX + E::NU; ++k) z[k] = __ldg(x + N * E::BLK + k);
                    } else {
#pragma unroll
                        for (int k = 0; k < E::CTRL_BLK; ++k) z[k] = __ldg(x + (N - 1) * E::BLK + E::NX + k);
#pragma unroll
                        for (int k = 0; k < E::NX; ++k) z[E::CTRL_BLK + k] = __ldg(x + N * E::BLK + k);
                    }
#pragma unroll
                    for (int k = 0; k < E::NV; ++k) z[E::FP_NLOC - E::NV + k] = __ldg(x + a.nvar - E::NV + k);
#pragma unroll
                    for (int r = 0; r < E::NP; ++r) lam[r] = (C::WH && y) ? __ldg(y + N * E::CB + r) : 0.0;
                    TimeArgs<E::NV> Tm;
                    time_args<E>(a, z + E::FP_NLOC - E::NV, p, N + 1, Tm);
                    EdgeElemSink<E, F, C::M_FP_J, C::M_FP_H, 0, false> s;
                    s.buf = buf + 2 * ne;
                    s.map = smap + a.fp_map_off;
                    s.crow = C::WC ? c + N * E::CB : nullptr;
                    s.red = red;
                    E::template fpath<C::WC, C::WJ, C::WH>(z, Tm, p, lam, s);
                }
            }
            if (warp_ == W_IP) {
                // implicit Euler: path constraints of node 1, g(t_1, x_1, U_1, v) (euler.jl:66-69: u(t_1) = U_1) -- the
                // step elements carry the path rows of nodes 2..N, the final-path element those of node N+1
                if constexpr (E::PATH_SHIFT && (C::WC || C::WJ || C::WH)) {
                    double z[E::FP_NLOC], lam[E::NP];
#pragma unroll
                    for (int k = 0; k < E::CTRL_BLK; ++k) z[k] = __ldg(x + E::NX + k);
#pragma unroll
                    for (int k = 0; k < E::NX; ++k) z[E::CTRL_BLK + k] = __ldg(x + k);
#pragma unroll
                    for (int k = 0; k < E::NV; ++k) z[E::FP_NLOC - E::NV + k] = __ldg(x + a.nvar - E::NV + k);
#pragma unroll
                    for (int r = 0; r < E::NP; ++r) lam[r] = (C::WH && y) ? __ldg(y + E::EQS + r) : 0.0;
                    TimeArgs<E::NV> Tm;
                    time_args<E>(a, z + E::FP_NLOC - E::NV, p, 1, Tm);
                    EdgeElemSink<E, F, C::M_FP_J, C::M_FP_H, 0, false> s;
                    s.buf = buf + 4 * ne;
                    s.map = smap + a.ip_map_off;
                    s.crow = C::WC ? c + E::EQS : nullptr;
                    s.red = red;
                    E::template fpath<C::WC, C::WJ, C::WH>(z, Tm, p, lam, s);
                }
            }
            if (warp_ == W_BC) {
                // boundary constraints + Mayer cost  (DOCP_functions.jl:103-111, 35-39)
                double z[E::BC_NLOC], lam[E::NB > 0 ? E::NB : 1];
                const long long N = a.N;
#pragma unroll
                for (int k = 0; k < E::NX; ++k) { z[k] = __ldg(x + k); z[E::NX + k] = __ldg(x + N * E::BLK + k); }
#pragma unroll
                for (int k = 0; k < E::NV; ++k) z[2 * E::NX + k] = __ldg(x + a.nvar - E::NV + k);
#pragma unroll
                for (int r = 0; r < E::NB; ++r) lam[r] = (C::WH && y) ? __ldg(y + a.ncon - E::NB + r) : 0.0;
                EdgeElemSink<E, F, C::M_BC_J, C::M_BC_H, C::M_BC_G, true> s;
                s.buf = buf + 3 * ne;
                s.map = smap + a.bc_map_off;
                s.crow = (C::WC && E::NB > 0) ? c + (a.ncon - E::NB) : nullptr;
                s.red = red;
                if constexpr (C::WC || C::WJ || C::WH) E::template bc<C::WC, C::WJ, C::WH>(z, p, lam, a.sigma, s);
                if constexpr (C::WG) E::grad_bc(z, p, s);
                if constexpr (C::WO) {
                    // a Mayer-only objective is this one number: no grid-wide sum
                    if constexpr (C::RO) red[C::R_O] += E::mayer(z, p);
                    else if (obj) *obj = E::mayer(z, p);
                }
            }
        }
    }
    __syncthreads();
    // ---- place the staged values of the step tasks (tasks 0/2 -> layer 0, task 1 -> layer 1: inside a
    //      layer no two values share an entry, so these are plain stores)
    for (int t = 0; t < a.ntask; ++t) {
        const int* mp = smap + a.task[t].map_off;
        const double* st = stage + t * C::M_STEP;
        double* layer = buf + (t & 1) * ne;
        auto place = [&](int lo, int hi) {
            for (int k = lo + tid; k < hi; k += T) {
                const int m = mp[k];
                if (m >= 0) layer[m] = st[k];
            }
        };
        if constexpr (C::WJ) place(C::M_JO, C::M_HO);
        if constexpr (C::WH) place(C::M_HO, C::M_GO);
        if constexpr (C::WG) place(C::M_GO, C::M_STEP);
    }
    __syncthreads();
    // ---- scatter the edge buffer (layers summed in a fixed order)
    auto total = [&](int k) {
        const double t4 = ((buf[k] + buf[ne + k]) + buf[2 * ne + k]) + buf[3 * ne + k];
        if constexpr (E::PATH_SHIFT) return t4 + buf[4 * ne + k];
        else return t4;
    };
    if constexpr (C::WJ)
        for (int k = tid; k < a.ne_j; k += T) jvals[spos[k]] = total(k);
    if constexpr (C::WH)
        for (int k = a.ne_j + tid; k < a.ne_j + a.ne_h; k += T) hvals[spos[k]] = total(k);
    if constexpr (C::WG)
        for (int k = a.ne_j + a.ne_h + tid; k < ne; k += T) grad[spos[k]] = total(k);
    __syncthreads();   // the caller reuses the shared memory of `buf` for its reduction scratch
}

// One phase of a warp tile: evaluate, stage, add the halo contributions, copy out.
template <class E, int F, bool PC, bool PJ, bool PH, bool PG, bool PO>
__device__ __forceinline__ void tile_phase(const EvalArgs& a, double* __restrict__ ws, const double* __restrict__ z,
                                           const double* __restrict__ lam, const TimeArgs<E::NV>& Tm,
                                           const double* __restrict__ p, double* red, int lane, int nown, int i0,
                                           bool active, double* c, double* jvals, double* hvals, double* grad) {
    using C = Cfg<E, F>;
    // global destinations of this tile and the matching shared-memory phases
    double* gC = PC ? c + (long long)(i0 - 1) * E::CB : nullptr;
    double* gJ = PJ ? jvals + (E::J_FIRST + (long long)(i0 - 2) * E::J_BLOCK) : nullptr;
    double* gH = PH ? hvals + (E::H_FIRST + (long long)(i0 - 2) * E::H_BLOCK) : nullptr;
    double* gG = PG ? grad + (long long)(i0 - 1) * E::BLK : nullptr;
    double* sC = ws + C::S_C + (PC ? phase_of(gC) : 0);
    double* sJ = ws + C::S_J + (PJ ? phase_of(gJ) : 0);
    double* sH = ws + C::S_H + (PH ? phase_of(gH) : 0);
    double* sG = ws + C::S_G + (PG ? phase_of(gG) : 0);
    double* gJV[C::NV > 0 ? C::NV : 1];
    double* sJV[C::NV > 0 ? C::NV : 1];
    if constexpr (PJ && C::NV > 0) {
#pragma unroll
        for (int k = 0; k < C::NV; ++k) {
            gJV[k] = jvals + (a.jv_base[k] + (long long)(i0 - 1) * E::JV_STRIDE);
            sJV[k] = ws + C::S_JV + k * C::JV_ONE + phase_of(gJV[k]);
        }
    }
    TileSink<E, F, PC, PJ, PH, PG> s;
    s.red = red;
    s.owner = active && lane >= 1;
    s.giver = active && lane < nown;
    const int ob = s.owner ? lane - 1 : C::OWN;   // own block (dump block for the halo lane)
    const int nb = s.giver ? lane : C::OWN;       // next block (dump block for the last lane)
    s.c_ = sC + ob * E::CB;
    s.cn_ = sC + nb * E::CB;
    s.jo_ = sJ + ob * C::JS;
    s.jn_ = sJ + nb * C::JS;
    s.ho_ = sH + ob * C::HS;
    s.go_ = sG + ob * E::BLK;
    if constexpr (PJ && C::NV > 0) {
#pragma unroll
        for (int k = 0; k < C::NV; ++k) {
            s.jv_[k] = sJV[k] + ob * E::JV_STRIDE;
            s.jvn_[k] = sJV[k] + nb * E::JV_STRIDE;
        }
    }
    if (active) {
        if constexpr (PG) {
#pragma unroll
            for (int k = 0; k < E::BLK; ++k) s.go_[k] = 0.0;
        }
        if constexpr (PC || PJ || PH) E::template step<PC, PJ, PH>(z, Tm, p, lam, a.sigma, s);
        if constexpr (PG) E::grad_step(z, Tm, p, s);
        if constexpr (PO && C::RO) { if (s.owner) red[C::R_O] += E::cost_step(z, Tm, p); }
    }
    __syncwarp();
    if constexpr ((PH && E::NH_NEXT > 0) || (PG && E::NG_NEXT > 0)) {
        if (active) {
            if constexpr (PH) {
#pragma unroll
                for (int n = 0; n < E::NH_NEXT; ++n) sH[nb * C::HS + E::H_NEXT(n, 2)] += s.hn_[n];
            }
            if constexpr (PG) {
#pragma unroll
                for (int n = 0; n < E::NG_NEXT; ++n) sG[nb * E::BLK + E::G_NEXT(n, 2)] += s.gn_[n];
            }
        }
        __syncwarp();
    }
    fence_proxy_async_smem();
    __syncwarp();
    if (nown > 0) {
        // the elected lane hands the 16-byte-aligned bodies of all ranges to the copy engine, lane 1 stores
        // the ragged ends (ranges start on 8-byte, not 16-byte, boundaries for odd block sizes)
        if (elect_one()) {
            if constexpr (PC) bulk_body(gC, smem_u32(sC), nown * E::CB);
            if constexpr (PJ) {
                if constexpr (!C::J_PADDED) bulk_body(gJ, smem_u32(sJ), nown * E::J_BLOCK);
                if constexpr (C::NV > 0) {
#pragma unroll
                    for (int k = 0; k < C::NV; ++k) bulk_body(gJV[k], smem_u32(sJV[k]), nown * E::JV_STRIDE);
                }
            }
            if constexpr (PH && !C::H_PADDED) bulk_body(gH, smem_u32(sH), nown * E::H_BLOCK);
            if constexpr (PG) bulk_body(gG, smem_u32(sG), nown * E::BLK);
            bulk_commit();
        }
        if constexpr ((PH && C::H_PADDED) || (PJ && C::J_PADDED)) {   // padded rows: lane l hands row l over (block size even: every row has the tile's phase)
            if (lane < nown) {
                if constexpr (PH && C::H_PADDED) {
                    bulk_body(gH + lane * E::H_BLOCK, smem_u32(sH + lane * C::HS), E::H_BLOCK);
                    ragged_ends(gH + lane * E::H_BLOCK, sH + lane * C::HS, E::H_BLOCK);
                }
                if constexpr (PJ && C::J_PADDED) {
                    bulk_body(gJ + lane * E::J_BLOCK, smem_u32(sJ + lane * C::JS), E::J_BLOCK);
                    ragged_ends(gJ + lane * E::J_BLOCK, sJ + lane * C::JS, E::J_BLOCK);
                }
                bulk_commit();
            }
        }
        if (lane == 1) {
            if constexpr (PC) ragged_ends(gC, sC, nown * E::CB);
            if constexpr (PJ) {
                if constexpr (!C::J_PADDED) ragged_ends(gJ, sJ, nown * E::J_BLOCK);
                if constexpr (C::NV > 0) {
#pragma unroll
                    for (int k = 0; k < C::NV; ++k) ragged_ends(gJV[k], sJV[k], nown * E::JV_STRIDE);
                }
            }
            if constexpr (PH && !C::H_PADDED) ragged_ends(gH, sH, nown * E::H_BLOCK);
            if constexpr (PG) ragged_ends(gG, sG, nown * E::BLK);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Deterministic reduction of the global sums ((v,v) Hessian entries, d cost / d v, cost, violation) over
// the grid -- only compiled for the slots a STEP contributes to (Cfg::RH / RG / RO / WV).  Inside a CTA:
// warp shuffles, then the last warp to arrive (shared-memory ticket, no CTA barrier) combines the CTA's NW
// warps in a fixed order into ONE partial per slot; across CTAs the partial is stored, fenced, a global
// ticket is taken and the last CTA of the instance sums all partials in a fixed order.
// (Measured and dropped: a collector warp in CTA 0 that polls self-validating 16-byte records instead of
// the fence + ticket -- 8.4 vs 7.3 us on bolza/midpoint at grid 1e5, 10.8 vs 8.3 us on the double integrator.)
template <class E, int F>
__device__ __forceinline__ void publish_reduction(const EvalArgs& a, const double* red, double* sred,
                                                  unsigned int* cta_ticket_p, long long inst, int cta, int lane,
                                                  int warp, double* hvals, double* grad, const int NW) {
    using C = Cfg<E, F>;
    if constexpr (C::NREDX > 0) {
        constexpr int NR = C::NREDX;
        // slot R_V (the constraint violation) is combined with max, every other slot with +
        auto comb = [](int k, double x, double y) { return (C::WV && k == C::R_V) ? nanmax(x, y) : x + y; };
        unsigned int& cta_ticket = *cta_ticket_p;
        const int nblk = a.ctas;
        double wv[NR];
#pragma unroll
        for (int k = 0; k < NR; ++k) {
            double v = red[k];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v = comb(k, v, __shfl_xor_sync(0xffffffffu, v, o));
            wv[k] = v;
        }
        unsigned int flag = 0;
        if (lane == 0) {
#pragma unroll
            for (int k = 0; k < NR; ++k) sred[warp * NR + k] = wv[k];
            __threadfence_block();
            flag = (atomicAdd(&cta_ticket, 1u) == (unsigned int)(NW - 1));
        }
        flag = __shfl_sync(0xffffffffu, flag, 0);
        if (!flag) return;
        __threadfence_block();
        double* part = a.partials + (inst * nblk) * NR;
        if (lane < NR) {
            double v = sred[lane];
            for (int w = 1; w < NW; ++w) v = comb(lane, v, sred[w * NR + lane]);
            part[(long long)cta * NR + lane] = v;
            __threadfence();
        }
        __syncwarp();
        unsigned int last = 0;
        if (lane == 0) last = (atomicAdd(&a.counters[inst], 1u) == (unsigned int)(nblk - 1));
        last = __shfl_sync(0xffffffffu, last, 0);
        if (!last) return;
        __threadfence();
        double acc[NR];
#pragma unroll
        for (int k = 0; k < NR; ++k) acc[k] = 0.0;
        // 16 independent loads per lane in flight (one memory round trip per 512 partials)
        for (int base = 0; base < nblk; base += 32 * 16) {
#pragma unroll
            for (int k = 0; k < NR; ++k) {
                double t[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int bb = base + j * 32 + lane;
                    t[j] = bb < nblk ? __ldcg(&part[(long long)bb * NR + k]) : 0.0;
                }
#pragma unroll
                for (int j = 0; j < 16; ++j) acc[k] = comb(k, acc[k], t[j]);
            }
        }
        double violv = 0.0;
#pragma unroll
        for (int k = 0; k < NR; ++k) {
            double v = acc[k];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v = comb(k, v, __shfl_xor_sync(0xffffffffu, v, o));
            if (lane == 0) {
                if (C::WV && k == C::R_V) { violv = v; if (a.viol) a.viol[inst] = v; }
                else if (C::RH && k < C::R_G) { if (a.hvv_pos[k - C::R_H] >= 0) hvals[a.hvv_pos[k - C::R_H]] = v; }
                else if (C::RG && k < C::R_O) grad[a.nvar - E::NV + (k - C::R_G)] = v;
                else if (C::RO && a.obj) a.obj[inst] = v;
            }
        }
        if (lane == 0) a.counters[inst] = 0u;
        if constexpr (C::WV) {
            if (a.peer_base && lane == 0) {
                // ---- fused gather: the two records of this instance go to every rank (see EvalArgs).  The objective was
                //      summed above or written by the boundary job of the instance's edge CTA, whose ticket this CTA has seen.
                const long long nb = (long long)gridDim.y;
                const unsigned long long ob = (unsigned long long)__double_as_longlong(a.obj ? __ldcg(a.obj + inst) : 0.0);
                const unsigned long long vb = (unsigned long long)__double_as_longlong(violv);
                for (int spin = 0; spin < (1 << 23); ++spin) {   // the gate of this launch (opened by its first CTA)
                    unsigned long long g;
                    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(g) : "l"(a.peer_gate) : "memory");
                    if (g >= a.peer_epoch) break;
                    __nanosleep(200);
                }
                for (int pr_ = 0; pr_ < a.peer_world; ++pr_) {
                    double* rec = a.peer_base[pr_] + a.peer_rec;
                    asm volatile("st.global.v2.u64 [%0], {%1, %2};" :: "l"(rec + 2 * inst), "l"(ob), "l"(ob ^ a.peer_key) : "memory");
                    asm volatile("st.global.v2.u64 [%0], {%1, %2};" :: "l"(rec + 2 * (nb + inst)), "l"(vb), "l"(vb ^ a.peer_key) : "memory");
                }
            }
        }
    }
}

template <class E, int F>
__device__ __forceinline__ void eval_body(const EvalArgs& a) {
    using C = Cfg<E, F>;
    extern __shared__ __align__(16) double smem[];
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int nw = (int)(blockDim.x >> 5);   // warps per CTA: chosen per launch (<= C::WPC)
    const int cta = blockIdx.x;
    const long long inst = blockIdx.y;

    const double* __restrict__ x = a.x + inst * a.x_stride;
    const double* __restrict__ y = a.y ? a.y + inst * a.y_stride : nullptr;
    double* c = C::WC ? a.c + inst * a.c_stride : nullptr;
    double* jvals = C::WJ ? a.jvals + inst * a.j_stride : nullptr;
    double* hvals = C::WH ? a.hvals + inst * a.h_stride : nullptr;
    double* grad = C::WG ? a.grad + inst * a.g_stride : nullptr;

    double red[C::NREDX > 0 ? C::NREDX : 1];
#pragma unroll
    for (int k = 0; k < (C::NREDX > 0 ? C::NREDX : 1); ++k) red[k] = 0.0;
    __shared__ unsigned int cta_ticket;
    if constexpr (C::NREDX > 0) {
        if (tid == 0) cta_ticket = 0u;
        __syncthreads();
    }
    // the parameters of the instance live in registers: both phases (and the compiler's common
    // subexpressions across them, e.g. Goddard's exp) see the same values instead of reloading them
    double pr[E::NPAR > 0 ? E::NPAR : 1];
    {
        const double* __restrict__ pg = a.params + inst * (E::NPAR > 0 ? E::NPAR : 1);
#pragma unroll
        for (int k = 0; k < E::NPAR; ++k) pr[k] = __ldg(pg + k);
    }

    CTD_STAMP(0);
    if (cta == 0) {
        edge_cta<E, F>(a, x, y, pr, c, jvals, hvals, grad, (C::WO && a.obj) ? a.obj + inst : nullptr, smem, red, nw);
        CTD_STAMP(4);
    } else {
        const int tile = (cta - 1) * nw + warp;
        const int i0 = 2 + tile * C::OWN;                             // first owned step / column block
        const int left = a.N - i0;                                    // owned blocks end with block N - 1
        const int nown = left < C::OWN ? (left < 0 ? 0 : left) : C::OWN;
        const int i = i0 - 1 + lane;                                  // this thread's step (lane 0: halo)
        const bool active = nown > 0 && lane <= nown;
        double* ws = smem + warp * C::WARP_DOUBLES;
        const double* __restrict__ xs = x + (long long)(i - 1) * E::BLK;
        const double* __restrict__ ys = (C::WH && y) ? y + (long long)(i - 1) * E::CB : nullptr;
        // ---- before the grid dependency: normalized times of the step, L2 prefetch of the tile's inputs
        double si = 0.0, sp = 0.0;
        if (active) {
            si = grid_point(a, i - 1);
            sp = grid_point(a, i);
            // one 128-byte line per lane and pass: lanes 0..15 walk the x range of the tile, 16..31 its y range
            // (every address lies inside the range: a prefetch must not touch what may be unmapped)
            const int half = lane & 15;
            if (lane < 16) {
                const int bytes = ((nown + 1) * E::BLK + E::NNEXT) * 8;
                const char* b0 = reinterpret_cast<const char*>(x + (long long)(i0 - 2) * E::BLK);
                for (int o = half * 128; o < bytes + 128; o += 16 * 128) prefetch_l2(b0 + (o < bytes - 8 ? o : bytes - 8));
            } else if (C::WH && y) {
                const int bytes = (nown + 1) * E::CB * 8;
                const char* b0 = reinterpret_cast<const char*>(y + (long long)(i0 - 2) * E::CB);
                for (int o = half * 128; o < bytes + 128; o += 16 * 128) prefetch_l2(b0 + (o < bytes - 8 ? o : bytes - 8));
            }
            if (lane == 0 && E::NV > 0) prefetch_l2(x + a.nvar - E::NV);
        }
        pdl_entry();
        CTD_STAMP(7);
        double z[E::NZ], lam[E::CB];
        TimeArgs<E::NV> Tm;
        if (active) {
#pragma unroll
            for (int k = 0; k < E::NOWN + E::NNEXT; ++k) z[k] = __ldg(xs + k);
#pragma unroll
            for (int k = 0; k < E::NV; ++k) z[E::NOWN + E::NNEXT + k] = __ldg(x + a.nvar - E::NV + k);
#pragma unroll
            for (int r = 0; r < E::CB; ++r) lam[r] = ys ? __ldg(ys + r + ((E::PATH_SHIFT && r >= E::EQS) ? E::CB : 0)) : 0.0;
            time_args_s<E>(si, sp, z + E::NOWN + E::NNEXT, pr, Tm);
        }
#ifdef CTD_TRACE
        if (active && Tm.hh == -1.2345) CTD_STAMP(5);   // forces the loads to have arrived
        CTD_STAMP(1);
#endif
        if constexpr (C::GZ) {   // gradient entries of the owned column blocks: zeros (see Cfg::GZ)
            if (nown > 0) {
                double* g0 = grad + (long long)(i0 - 1) * E::BLK;
                for (int k = lane; k < nown * E::BLK; k += 32) g0[k] = 0.0;
            }
        }
        // (one pass over the step code for both phases -- the sink switching phase at the first Hessian value, so that
        //  exp / reciprocals / midpoints are computed once -- was measured: 10.02 vs 9.95 us, the kernel is not bound by
        //  instruction issue)
        if constexpr (C::PHASE_A)
            tile_phase<E, F, C::WC, C::WJ, false, C::SG, C::WO>(a, ws, z, lam, Tm, pr, red, lane, nown, i0, active, c,
                                                                  jvals, hvals, grad);
        CTD_STAMP(2);
        if constexpr (C::PHASE_B) {
            if constexpr (C::PHASE_A) {   // phase B reuses the region: the copy engine must have read it
                bulk_wait_read();         // (returns at once in the lanes that issued no copy)
                __syncwarp();
            }
            CTD_STAMP(3);
            tile_phase<E, F, false, false, true, false, false>(a, ws, z, lam, Tm, pr, red, lane, nown, i0, active, c,
                                                                jvals, hvals, grad);
        }
        CTD_STAMP(4);
    }

    // the partial sums are published while the copy engine is still draining the last phase
    publish_reduction<E, F>(a, red, smem + nw * C::WARP_DOUBLES, &cta_ticket, inst, cta, lane, warp, hvals, grad, nw);
    CTD_STAMP(5);
    if (cta != 0) bulk_wait_read();   // shared memory must outlive the copies that read it (no-op for lanes without copies)
    CTD_STAMP(6);
    CTD_STAMP_SMID();
}

// =============================================================================================
// "Parts" kernel for wide elements on small grids (e.g. 3-stage IRK at N = 2e4: only 645 tiles, four
// warps per SM if one warp did a whole tile).  A CTA of P warps owns ONE tile; every warp evaluates the
// same 31+1 steps but keeps only its share of the outputs (the generated straight-line code is
// instantiated once per part; the unused values and everything that only feeds them are removed by
// dead-code elimination).  All outputs of the tile are staged together; after a CTA barrier warp 0
// hands the ranges to the copy engine.
// =============================================================================================
#ifndef CTD_PARTS
#define CTD_PARTS 2
#endif
#ifndef CTD_PART_HW
#define CTD_PART_HW 1
#endif
#ifndef CTD_PART_OWN
#define CTD_PART_OWN 31
#endif

#ifndef CTD_PARTS_TWO_PHASE
#define CTD_PARTS_TWO_PHASE 1
#endif


template <class E, int F>
struct PCfg {
    using C = Cfg<E, F>;
    static constexpr int P = CTD_PARTS;
    static constexpr int T = 32 * P;
    // two phases like the one-warp kernel (A: constraints + Jacobian + gradient, B: Hessian, same shared
    // memory) when both are requested: a 3-stage IRK step stages 290 doubles, 186 at a time with two phases
    static constexpr bool TWO = CTD_PARTS_TWO_PHASE && C::WJ && C::WH;
    static constexpr int PER_A = (C::WC ? E::CB : 0) + (C::WJ ? C::JS + C::NJV : 0) + (C::WG ? E::BLK : 0);
    static constexpr int PER_B = C::WH ? C::HS : 0;
    static constexpr int PER_STEP = TWO ? (PER_A > PER_B ? PER_A : PER_B) : PER_A + PER_B;
    static constexpr int FIT = PER_STEP > 0 ? (C::SMEM_MAX - 2048) / (PER_STEP * 8) - 1 : 31;
    // tiles of 24..30 steps when that lets 5 CTAs share an SM (44 KB each) instead of 4
    static constexpr int FIT5 = PER_STEP > 0 ? (44 * 1024) / (PER_STEP * 8) - 1 : 31;
    static constexpr bool FIVE = TWO && FIT5 >= 24 && FIT5 < CTD_PART_OWN;
    static constexpr int OWN_T = FIVE ? FIT5 : CTD_PART_OWN;
    static constexpr int OWN = (OWN_T < FIT) ? OWN_T : FIT;
    // The single-output kernels of an element whose fused kernel runs with 5 CTAs per SM get the same register
    // cap: the Hessian-only kernel of a 3-stage IRK otherwise takes 248 registers = 4 CTAs per SM = 592 of the
    // 645 tiles of a 2e4-step grid resident, i.e. a second wave (the same code is spill-free in 168 registers
    // inside the two-phase kernel).  Elements outside that class keep the full register file.
    static constexpr int PER7 = (E::CB + C::JS + C::NJV) > C::HS ? (E::CB + C::JS + C::NJV) : C::HS;
    static constexpr int FIT5_7 = (44 * 1024) / (PER7 * 8) - 1;
    static constexpr bool FIVE7 = FIT5_7 >= 24 && FIT5_7 < CTD_PART_OWN;
    static constexpr int MINB = (FIVE || (FIVE7 && PER_STEP * 8 * (OWN + 1) <= 44 * 1024)) ? 5 : 1;
    static_assert(OWN >= 1, "element too wide for one SM's shared memory");
    static constexpr int even(int n) { return (n + 3) & ~1; }
    static constexpr int S_C = 0;
    static constexpr int S_J = S_C + (C::WC ? even((OWN + 1) * E::CB) : 0);
    static constexpr int S_JV = S_J + (C::WJ ? even((OWN + 1) * C::JS) : 0);
    static constexpr int JV_ONE = even((OWN + 1) * E::JV_STRIDE);
    static constexpr int A_END0 = S_JV + (C::WJ ? C::NV * JV_ONE : 0);
    static constexpr int H_LEN = C::WH ? even((OWN + 1) * C::HS) : 0;
    static constexpr int G_LEN = C::WG ? even((OWN + 1) * E::BLK) : 0;
    // one phase: c | J | Jv | H | g     two phases: A = c | J | Jv | g, B = H (from 0)
    static constexpr int S_H = TWO ? 0 : A_END0;
    static constexpr int S_G = TWO ? A_END0 : A_END0 + H_LEN;
    static constexpr int S_OUT = TWO ? ((A_END0 + G_LEN) > H_LEN ? (A_END0 + G_LEN) : H_LEN) : A_END0 + H_LEN + G_LEN;
    // the tile's inputs are staged at the start of the same region before any output is (see the kernel)
    static constexpr int S_IN = even((OWN + 1) * E::BLK + E::NNEXT) + even((OWN + 1) * E::CB);
    static constexpr int S_RED0 = S_OUT > S_IN ? S_OUT : S_IN;
    static constexpr int S_RED = (S_RED0 + 1) & ~1;
    static constexpr int S_END = S_RED + P * (C::NREDX > 0 ? C::NREDX : 1) + 2;
    static constexpr size_t TILE_SMEM = (size_t)S_END * sizeof(double);
    // every output of the element gets a running index
    static constexpr int O_C = 0;
    static constexpr int O_JO = O_C + (C::WC ? E::CB : 0);
    static constexpr int O_JN = O_JO + (C::WJ ? E::NJ_OWN : 0);
    static constexpr int O_JV = O_JN + (C::WJ ? E::NJ_NEXT : 0);
    static constexpr int O_HO = O_JV + (C::WJ ? E::NJ_V : 0);
    static constexpr int O_HN = O_HO + (C::WH ? E::NH_OWN : 0);
    static constexpr int O_HV = O_HN + (C::WH ? E::NH_NEXT : 0);
    static constexpr int O_GO = O_HV + (C::WH ? C::NVV : 0);
    static constexpr int O_GN = O_GO + (C::WG ? E::NG_OWN : 0);
    static constexpr int O_GV = O_GN + (C::WG ? E::NG_NEXT : 0);
    static constexpr int TOT = O_GV + (C::WG ? E::NG_V : 0);
    // a Hessian value costs about CTD_PART_HW times the instructions of a Jacobian value: the one-phase
    // split is by weighted index so that the parts take about the same time
    static constexpr long long wpos(int g) { return g < O_HO ? g : (g < O_GO ? O_HO + (long long)(g - O_HO) * CTD_PART_HW : O_HO + (long long)(O_GO - O_HO) * CTD_PART_HW + (g - O_GO)); }
    // two phases: the outputs of each phase are split evenly on their own
    static constexpr int NA = O_HO + (TOT - O_GO), NB = O_GO - O_HO;
    static constexpr int part_of(int g) {
        if (TOT == 0) return 0;
        if (!TWO) return (int)(wpos(g) * P / wpos(TOT));
        const bool hess = g >= O_HO && g < O_GO;
        return hess ? (int)((long long)(g - O_HO) * P / NB) : (int)((long long)(g < O_HO ? g : O_HO + (g - O_GO)) * P / NA);
    }
};

// PA / PB: which outputs this pass produces (both in the one-phase kernel)
template <class E, int F, int PART, bool PA, bool PB>
struct PartSink {
    using C = Cfg<E, F>;
    using PC = PCfg<E, F>;
    static constexpr bool WC = PA && C::WC, WJ = PA && C::WJ, WG = PA && C::WG, WH = PB && C::WH;
    double *c_, *jo_, *jn_, *ho_, *go_;
    double* jv_[E::NV > 0 ? E::NV : 1];
    bool owner;
    double hn_[E::NH_NEXT > 0 ? E::NH_NEXT : 1];
    double gn_[E::NG_NEXT > 0 ? E::NG_NEXT : 1];
    double* red;
    static constexpr bool mine(int g) { return PC::part_of(g) == PART; }
    template <int R> __device__ __forceinline__ void c(double v) {
        if constexpr (WC && mine(PC::O_C + R)) {
            c_[R] = v;
            if constexpr (C::WV) { if (owner) red[C::R_V] = nanmax(red[C::R_V], row_violation(E::CON_LB(R), E::CON_UB(R), v)); }
        }
    }
    template <int N> __device__ __forceinline__ void jo(double v) { if constexpr (WJ && mine(PC::O_JO + N)) jo_[E::J_OWN(N, 2)] = v; }
    template <int N> __device__ __forceinline__ void jn(double v) { if constexpr (WJ && mine(PC::O_JN + N)) jn_[E::J_NEXT(N, 2)] = v; }
    template <int N> __device__ __forceinline__ void jv(double v) {
        if constexpr (WJ && mine(PC::O_JV + N)) jv_[E::J_V(N, 1)][E::J_V(N, 2)] = v;
    }
    template <int N> __device__ __forceinline__ void ho(double v) { if constexpr (WH && mine(PC::O_HO + N)) ho_[E::H_OWN(N, 2)] = v; }
    template <int N> __device__ __forceinline__ void hn(double v) { if constexpr (WH && mine(PC::O_HN + N)) hn_[N] = v; }
    template <int K> __device__ __forceinline__ void hv(double v) {
        if constexpr (WH && C::RH && mine(PC::O_HV + K)) red[C::R_H + K] += owner ? v : 0.0;
    }
    template <int N> __device__ __forceinline__ void go(double v) { if constexpr (WG && mine(PC::O_GO + N)) go_[E::G_OWN(N, 2)] = v; }
    template <int N> __device__ __forceinline__ void gn(double v) { if constexpr (WG && mine(PC::O_GN + N)) gn_[N] = v; }
    template <int N> __device__ __forceinline__ void gv(double v) {
        if constexpr (WG && C::RG && mine(PC::O_GV + N)) red[C::R_G + E::G_V(N, 2)] += owner ? v : 0.0;
    }
};

template <class E, int F, int PART, bool PA, bool PB>
__device__ __forceinline__ void run_part(const EvalArgs& a, const double* z, const double* lam, const TimeArgs<E::NV>& Tm,
                                         const double* __restrict__ p, double* red, bool active, bool owner, int ob, int nb,
                                         double* sC, double* sJ, double* const* sJV, double* sH, double* sG) {
    using C = Cfg<E, F>;
    using PC = PCfg<E, F>;
    using S = PartSink<E, F, PART, PA, PB>;
    S s;
    s.red = red;
    s.owner = owner;
    s.c_ = sC + ob * E::CB;
    s.jo_ = sJ + ob * C::JS;
    s.jn_ = sJ + nb * C::JS;
    s.ho_ = sH + ob * C::HS;
    s.go_ = sG + ob * E::BLK;
    if constexpr (S::WJ && C::NV > 0) {
#pragma unroll
        for (int k = 0; k < C::NV; ++k) s.jv_[k] = sJV[k] + ob * E::JV_STRIDE;
    }
    if (active) {
        if constexpr (S::WG && !E::LAGRANGE && PART == 0) {
#pragma unroll
            for (int k = 0; k < E::BLK; ++k) s.go_[k] = 0.0;
        }
        if constexpr (S::WC || S::WJ || S::WH) E::template step<S::WC, S::WJ, S::WH>(z, Tm, p, lam, a.sigma, s);
        if constexpr (S::WG) E::grad_step(z, Tm, p, s);
        if constexpr (PA && C::RO && PART == 0) { if (owner) red[C::R_O] += E::cost_step(z, Tm, p); }
    }
    __syncthreads();   // all owners have stored: the neighbours' contributions can be added
    if (active) {
        if constexpr (S::WH) {
#pragma unroll
            for (int n = 0; n < E::NH_NEXT; ++n)
                if (S::mine(PC::O_HN + n)) sH[nb * C::HS + E::H_NEXT(n, 2)] += s.hn_[n];
        }
        if constexpr (S::WG) {
#pragma unroll
            for (int n = 0; n < E::NG_NEXT; ++n)
                if (S::mine(PC::O_GN + n)) sG[nb * E::BLK + E::G_NEXT(n, 2)] += s.gn_[n];
        }
    }
}

template <class E, int F, int PART, bool PA, bool PB>
struct PartDispatch {
    template <class... A>
    __device__ __forceinline__ static void run(int warp, A&&... args) {
        if (warp == PART) run_part<E, F, PART, PA, PB>(args...);
        else if constexpr (PART + 1 < PCfg<E, F>::P) PartDispatch<E, F, PART + 1, PA, PB>::run(warp, args...);
    }
};

template <class E, int F>
__global__ void __launch_bounds__(PCfg<E, F>::T, PCfg<E, F>::MINB) ctd_eval_kernel_parts(const EvalArgs a) {
    static_assert(!E::PATH_SHIFT, "shifted path rows (implicit Euler) are handled by the one-warp kernel only");
    using C = Cfg<E, F>;
    using PC = PCfg<E, F>;
    extern __shared__ __align__(16) double smem[];
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int cta = blockIdx.x;
    const long long inst = blockIdx.y;
    const double* __restrict__ x = a.x + inst * a.x_stride;
    const double* __restrict__ y = a.y ? a.y + inst * a.y_stride : nullptr;
    const double* __restrict__ p = a.params + inst * (E::NPAR > 0 ? E::NPAR : 1);
    double* c = C::WC ? a.c + inst * a.c_stride : nullptr;
    double* jvals = C::WJ ? a.jvals + inst * a.j_stride : nullptr;
    double* hvals = C::WH ? a.hvals + inst * a.h_stride : nullptr;
    double* grad = C::WG ? a.grad + inst * a.g_stride : nullptr;
    double red[C::NREDX > 0 ? C::NREDX : 1];
#pragma unroll
    for (int k = 0; k < (C::NREDX > 0 ? C::NREDX : 1); ++k) red[k] = 0.0;
    __shared__ unsigned int cta_ticket;
    if constexpr (C::NREDX > 0) {
        if (tid == 0) cta_ticket = 0u;
        __syncthreads();
    }
    CTD_STAMP(0);
    if (cta == 0) {
        edge_cta<E, F>(a, x, y, p, c, jvals, hvals, grad, (C::WO && a.obj) ? a.obj + inst : nullptr, smem, red, PC::P);
        CTD_STAMP(4);
    } else {
        const long long i0 = 2 + (long long)(cta - 1) * PC::OWN;
        const long long left = (long long)a.N - 1 - i0 + 1;
        const int nown = (int)(left < PC::OWN ? (left < 0 ? 0 : left) : PC::OWN);
        const long long i = i0 - 1 + lane;
        const bool active = nown > 0 && lane <= nown;
        // ---- the tile's inputs: the x blocks of steps i0-1 .. i0-1+nown (+ the head of the next block) and the matching
        //      rows of y are two contiguous ranges.  They are prefetched into L2 before the grid dependency is resolved and
        //      then staged in shared memory with coalesced loads by the whole CTA: a wide step block read per lane touches
        //      one 128-byte line per lane and per value (3-stage IRK: 13 + 12 values per step, ~700 L1 wavefronts per warp,
        //      2 us of the tile's 12 us); staged, each lane picks its values from shared memory without bank conflicts
        //      for odd block sizes.  (The staging area is the start of the output region, overwritten after a barrier.)
        const int nxt = nown > 0 ? (nown + 1) * E::BLK + E::NNEXT : 0;
        const int nyt = (nown > 0 && C::WH && y) ? (nown + 1) * E::CB : 0;
        const double* __restrict__ xb = x + (i0 - 2) * E::BLK;
        const double* __restrict__ yb = y ? y + (i0 - 2) * E::CB : nullptr;
        for (int o = tid * 128; o < nxt * 8 + 128; o += PC::T * 128)
            if (nxt > 0) prefetch_l2(reinterpret_cast<const char*>(xb) + (o < nxt * 8 - 8 ? o : nxt * 8 - 8));
        for (int o = tid * 128; o < nyt * 8 + 128; o += PC::T * 128)
            if (nyt > 0) prefetch_l2(reinterpret_cast<const char*>(yb) + (o < nyt * 8 - 8 ? o : nyt * 8 - 8));
        double si = 0.0, sp = 0.0;   // normalized times of this thread's step (immutable: read before the grid dependency)
        if (active) { si = grid_point(a, i - 1); sp = grid_point(a, i); }
        pdl_entry();
        CTD_STAMP(7);
        double* sx = smem;
        double* sy = smem + ((nxt + 1) & ~1);
        for (int k = tid; k < nxt; k += PC::T) sx[k] = __ldg(xb + k);
        for (int k = tid; k < nyt; k += PC::T) sy[k] = __ldg(yb + k);
        double vv[E::NV > 0 ? E::NV : 1];
#pragma unroll
        for (int k = 0; k < E::NV; ++k) vv[k] = __ldg(x + a.nvar - E::NV + k);
        __syncthreads();
        const bool owner = active && lane >= 1;
        const bool giver = active && lane < nown;
        const int ob = owner ? lane - 1 : PC::OWN;
        const int nb = giver ? lane : PC::OWN;
        double* gC = C::WC ? c + (i0 - 1) * E::CB : nullptr;
        double* gJ = C::WJ ? jvals + E::J_FIRST + (i0 - 2) * E::J_BLOCK : nullptr;
        double* gH = C::WH ? hvals + E::H_FIRST + (i0 - 2) * E::H_BLOCK : nullptr;
        double* gG = C::WG ? grad + (i0 - 1) * E::BLK : nullptr;
        double* sC = smem + PC::S_C + (C::WC ? phase_of(gC) : 0);
        double* sJ = smem + PC::S_J + (C::WJ ? phase_of(gJ) : 0);
        double* sH = smem + PC::S_H + (C::WH ? phase_of(gH) : 0);
        double* sG = smem + PC::S_G + (C::WG ? phase_of(gG) : 0);
        double* gJV[C::NV > 0 ? C::NV : 1];
        double* sJV[C::NV > 0 ? C::NV : 1];
        if constexpr (C::WJ && C::NV > 0) {
#pragma unroll
            for (int k = 0; k < C::NV; ++k) {
                gJV[k] = jvals + a.jv_base[k] + (i0 - 1) * E::JV_STRIDE;
                sJV[k] = smem + PC::S_JV + k * PC::JV_ONE + phase_of(gJV[k]);
            }
        }
        double z[E::NZ], lam[E::CB];
        TimeArgs<E::NV> Tm;
        if (active) {
#pragma unroll
            for (int k = 0; k < E::NOWN + E::NNEXT; ++k) z[k] = sx[lane * E::BLK + k];
#pragma unroll
            for (int k = 0; k < E::NV; ++k) z[E::NOWN + E::NNEXT + k] = vv[k];
#pragma unroll
            for (int r = 0; r < E::CB; ++r) lam[r] = nyt > 0 ? sy[lane * E::CB + r] : 0.0;
            time_args_s<E>(si, sp, z + E::NOWN + E::NNEXT, p, Tm);
        }
        __syncthreads();   // the staging area is the output region: every warp has its inputs in registers
#ifdef CTD_TRACE
        if (active && Tm.hh == -1.2345) CTD_STAMP(5);   // forces the loads to have arrived
        CTD_STAMP(1);
#endif
        // hands the staged ranges of phase A and / or B to the copy engine (warp 0: the elected lane issues the
        // bulk copies, lane 1 stores the ragged ends)
        auto store = [&](bool pa, bool pb) {
            if (warp == 0 && nown > 0) {
                const bool leader = elect_one();
                auto range = [&](double* g, double* s, int n) {
                    if (leader) bulk_body(g, smem_u32(s), n);
                    if (lane == 1) ragged_ends(g, s, n);
                };
                if (pa) {
                    if constexpr (C::WC) range(gC, sC, nown * E::CB);
                    if constexpr (C::WJ) {
                        if constexpr (!C::J_PADDED) range(gJ, sJ, nown * E::J_BLOCK);
                        if constexpr (C::NV > 0) {
#pragma unroll
                            for (int k = 0; k < C::NV; ++k) range(gJV[k], sJV[k], nown * E::JV_STRIDE);
                        }
                    }
                    if constexpr (C::WG) range(gG, sG, nown * E::BLK);
                }
                if (pb) {
                    if constexpr (C::WH && !C::H_PADDED) range(gH, sH, nown * E::H_BLOCK);
                }
                if (leader) bulk_commit();
                if constexpr ((C::WH && C::H_PADDED) || (C::WJ && C::J_PADDED)) {   // padded rows: lane l hands row l over (see Cfg::HS)
                    if (lane < nown) {
                        if constexpr (C::WH && C::H_PADDED) {
                            if (pb) {
                                bulk_body(gH + lane * E::H_BLOCK, smem_u32(sH + lane * C::HS), E::H_BLOCK);
                                ragged_ends(gH + lane * E::H_BLOCK, sH + lane * C::HS, E::H_BLOCK);
                            }
                        }
                        if constexpr (C::WJ && C::J_PADDED) {
                            if (pa) {
                                bulk_body(gJ + lane * E::J_BLOCK, smem_u32(sJ + lane * C::JS), E::J_BLOCK);
                                ragged_ends(gJ + lane * E::J_BLOCK, sJ + lane * C::JS, E::J_BLOCK);
                            }
                        }
                        bulk_commit();
                    }
                }
            }
        };
        if constexpr (!PC::TWO) {
            PartDispatch<E, F, 0, true, true>::run(warp, a, z, lam, Tm, p, red, active, owner, ob, nb, sC, sJ, sJV, sH, sG);
            CTD_STAMP(2);
            fence_proxy_async_smem();
            __syncthreads();
            CTD_STAMP(3);
            store(true, true);
        } else {
            // (every other CTA running the Hessian phase first, so that the stores of one half overlap the arithmetic of
            //  the other, was measured: 21.1 vs 15.1 us -- two code streams per SM)
            PartDispatch<E, F, 0, true, false>::run(warp, a, z, lam, Tm, p, red, active, owner, ob, nb, sC, sJ, sJV, sH, sG);
            fence_proxy_async_smem();
            __syncthreads();
            store(true, false);
            CTD_STAMP(2);
            // phase B reuses the region: the copy engine must have read it
            if (warp == 0 && nown > 0) bulk_wait_read();
            __syncthreads();
            CTD_STAMP(3);
            PartDispatch<E, F, 0, false, true>::run(warp, a, z, lam, Tm, p, red, active, owner, ob, nb, sC, sJ, sJV, sH, sG);
            fence_proxy_async_smem();
            __syncthreads();
            store(false, true);
        }
        if (warp == 0 && nown > 0) bulk_wait_read();   // shared memory must outlive the copies that read it
        CTD_STAMP(4);
    }
    publish_reduction<E, F>(a, red, smem + PC::S_RED, &cta_ticket, inst, cta, lane, warp, hvals, grad, PC::P);
    CTD_STAMP(6);
    CTD_STAMP_SMID();
}

// ---------------------------------------------------------------------------------------------
// Solution unpacking: getter() of src/ode/common.jl:7-104 with the normalisations
// of build_OCP_solution (src/DOCP_data.jl:514-633).  One thread per time node; outputs are
// [node][component] (the memory of the getter's matrices), one instance after the other.
struct UnpackArgs {
    const double* x;
    const double* y;          // may be nullptr: primal quantities only
    double *time, *state, *control, *variable, *costate, *path_dual, *boundary_dual;   // any may be nullptr
    const double* sgrid;
    const double* params;
    long long nvar, ncon;
    int N;
};

template <class E>
__global__ void __launch_bounds__(128) ctd_unpack_kernel(const UnpackArgs a) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x + 1;   // node 1..N+1
    const long long inst = blockIdx.y;
    const long long N = a.N;
    if (i > N + 1) return;
    const double* __restrict__ x = a.x + inst * a.nvar;
    const double* __restrict__ y = a.y ? a.y + inst * a.ncon : nullptr;
    const double* __restrict__ p = a.params + inst * (E::NPAR > 0 ? E::NPAR : 1);
    double v[E::NV > 0 ? E::NV : 1];
#pragma unroll
    for (int k = 0; k < E::NV; ++k) v[k] = __ldg(x + a.nvar - E::NV + k);
    const double t0 = E::T0_KIND == 0 ? E::T0_CONST : (E::T0_KIND == 1 ? v[E::T0_IDX > 0 ? E::T0_IDX : 0] : p[E::T0_IDX > 0 ? E::T0_IDX : 0]);
    const double tf = E::TF_KIND == 0 ? E::TF_CONST : (E::TF_KIND == 1 ? v[E::TF_IDX > 0 ? E::TF_IDX : 0] : p[E::TF_IDX > 0 ? E::TF_IDX : 0]);
    auto tgrid = [&](long long node) { return CTD_ADD(t0, CTD_MUL(__ldg(a.sgrid + node - 1), CTD_SUB(tf, t0))); };
    if (a.time) a.time[inst * (N + 1) + i - 1] = tgrid(i);
    if (a.state) {
#pragma unroll
        for (int k = 0; k < E::NX; ++k) a.state[(inst * (N + 1) + i - 1) * E::NX + k] = __ldg(x + (i - 1) * E::BLK + k);
    }
    if constexpr (E::NU > 0 && E::CONTROL_STEPS > 1) {
        // N * control_steps + 1 rows: every control of every step, then u(tf) = the FIRST control of the last step
        // (getter(:control), common.jl:82-99 with get_OCP_control_at_time_step(data, docp, N+1), l.140-155)
        if (a.control) {
            double* out = a.control + inst * (N * E::CONTROL_STEPS + 1) * E::NU;
            if (i <= N) {
#pragma unroll
                for (int k = 0; k < E::CONTROL_STEPS * E::NU; ++k) out[(i - 1) * E::CONTROL_STEPS * E::NU + k] = __ldg(x + (i - 1) * E::BLK + E::NX + k);
            } else {
#pragma unroll
                for (int k = 0; k < E::NU; ++k) out[N * E::CONTROL_STEPS * E::NU + k] = __ldg(x + (N - 1) * E::BLK + E::NX + k);
            }
        }
    } else if constexpr (E::NU > 0) {
        if (a.control) {
            double* out = a.control + (inst * (N + 1) + i - 1) * E::NU;
            if constexpr (E::STAGEWISE) {
                // b-weighted average of the stage controls, accumulated left to right (irk_stagewise.jl:197-205)
                const long long ci = (i == N + 1) ? N : i;
                const double bw[3] = {E::BB(0), E::BB(1), E::BB(2)};
#pragma unroll
                for (int k = 0; k < E::NU; ++k) {
                    double u = CTD_MUL(bw[0], __ldg(x + (ci - 1) * E::BLK + E::NX + k));
#pragma unroll
                    for (int j = 1; j < E::STAGE; ++j) u = CTD_ADD(u, CTD_MUL(bw[j], __ldg(x + (ci - 1) * E::BLK + E::NX + j * E::NU + k)));
                    out[k] = u;
                }
            } else {
                long long off;
                if (E::EULER_IMPLICIT) off = ((i == 1 ? 2 : i) - 2) * E::BLK + E::NX;      // euler.jl:66-69
                else off = (((i == N + 1) && !E::FINAL_CONTROL ? N : i) - 1) * E::BLK + E::NX;   // common.jl:140-155
#pragma unroll
                for (int k = 0; k < E::NU; ++k) out[k] = __ldg(x + off + k);
            }
        }
    }
    if (i == 1 && a.variable) {
#pragma unroll
        for (int k = 0; k < E::NV; ++k) a.variable[inst * E::NV + k] = v[k];
    }
    if (!y) return;
    if (a.costate && i <= N) {
#pragma unroll
        for (int k = 0; k < E::NX; ++k) a.costate[(inst * N + i - 1) * E::NX + k] = __ldg(y + (i - 1) * E::CB + k);
    }
    if constexpr (E::NP > 0) {
        if (a.path_dual) {
            const long long off = i <= N ? (i - 1) * E::CB + E::EQS : N * E::CB;
            const double h = i <= N ? CTD_SUB(tgrid(i + 1), tgrid(i)) : CTD_SUB(tgrid(N + 1), tgrid(N));
#pragma unroll
            for (int r = 0; r < E::NP; ++r) a.path_dual[(inst * (N + 1) + i - 1) * E::NP + r] = __ldg(y + off + r) / h;
        }
    }
    if (i == 1 && a.boundary_dual) {
#pragma unroll
        for (int r = 0; r < E::NB; ++r) a.boundary_dual[inst * E::NB + r] = __ldg(y + a.ncon - E::NB + r);
    }
}

// Initial guess / warm start (the inverse of unpacking): __initial_guess of src/DOCP_variables.jl:122-145 and
// irk_stagewise.jl:302-335 from SAMPLED state / control / variable data -- 0.1 everywhere, then the variables,
// the states at the nodes and the controls through the setters of src/ode/common.jl:184-225.
struct PackArgs {
    const double *state, *control, *variable, *stage_control;   // any may be nullptr (the 0.1 default stays)
    double* x;
    long long nvar;
    int N;
};

template <class E>
__global__ void __launch_bounds__(128) ctd_pack_kernel(const PackArgs a) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x + 1;   // node 1..N+1
    const long long inst = blockIdx.y;
    const long long N = a.N;
    if (i > N + 1) return;
    double* x = a.x + inst * a.nvar;
    const int blk = (i <= N) ? E::BLK : (int)(a.nvar - E::NV - N * E::BLK);     // last block: x_f (+ u_f for trapeze)
    double* xb = x + (i - 1) * E::BLK;
    for (int k = 0; k < blk; ++k) xb[k] = 0.1;
    if (a.state) {
#pragma unroll
        for (int k = 0; k < E::NX; ++k) xb[k] = a.state[(inst * (N + 1) + i - 1) * E::NX + k];
    }
    if constexpr (E::NU > 0) {
        if constexpr (E::STAGEWISE) {
            if (i <= N) {
#pragma unroll
                for (int j = 0; j < E::STAGE; ++j)
#pragma unroll
                    for (int k = 0; k < E::NU; ++k) {
                        if (a.stage_control) xb[E::NX + j * E::NU + k] = a.stage_control[((inst * N + i - 1) * E::STAGE + j) * E::NU + k];
                        else if (a.control) xb[E::NX + j * E::NU + k] = a.control[(inst * (N + 1) + i - 1) * E::NU + k];
                    }
            }
        } else if (a.control && (i <= N || E::FINAL_CONTROL)) {
            // the node's control for every control step of the time step (DOCP_variables.jl:138-140)
#pragma unroll
            for (int j = 0; j < E::CONTROL_STEPS; ++j)
#pragma unroll
                for (int k = 0; k < E::NU; ++k) xb[E::NX + j * E::NU + k] = a.control[(inst * (N + 1) + i - 1) * E::NU + k];
        }
    }
    if (i == 1) {
#pragma unroll
        for (int k = 0; k < E::NV; ++k) x[a.nvar - E::NV + k] = a.variable ? a.variable[inst * E::NV + k] : 0.1;
    }
}

// the same body under two register budgets (see Cfg::LIGHT)
template <class E, int F>
__global__ void __maxnreg__(CTD_MAXNREG) ctd_eval_kernel_light(const EvalArgs a) { eval_body<E, F>(a); }
template <class E, int F>
__global__ void __launch_bounds__(Cfg<E, F>::T) ctd_eval_kernel(const EvalArgs a) { eval_body<E, F>(a); }

}  // namespace ctd
