// Shared declarations between the generated element structs (csrc/generated/*.cuh), the kernels
// (ctd_kernels.cuh) and the host model (ctd_model.hpp).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#define CTD_INF (__builtin_huge_val())

// Non-contracting arithmetic for the scheme formulas (time grid, midpoint / stage points, defects):
// the reference's Julia code rounds every product and sum separately, whereas nvcc would fuse a*b+c
// into one FMA.  Feeding the nonlinear OCP functions bit-identical arguments keeps ill-conditioned
// derivative entries (e.g. Goddard's exp(-500 (r-1)) terms that cancel between stages) within 1e-12.
#ifdef __CUDA_ARCH__
#define CTD_MUL(a, b) __dmul_rn((a), (b))
#define CTD_ADD(a, b) __dadd_rn((a), (b))
#define CTD_SUB(a, b) __dsub_rn((a), (b))
#else
#define CTD_MUL(a, b) ((a) * (b))
#define CTD_ADD(a, b) ((a) + (b))
#define CTD_SUB(a, b) ((a) - (b))
#endif

namespace ctd {

// Times of one step, computed like the reference (t = t0 + s*(tf - t0), DOCP_data.jl:456;
// h = t_{i+1} - t_i) plus the forward-mode seeds of those quantities w.r.t. the free times in v.
template <int NV>
struct TimeArgs {
    double ti, tp, hh;
    double dti[NV > 0 ? NV : 1], dtp[NV > 0 ? NV : 1], dh[NV > 0 ? NV : 1];
};

struct Box {  // (lb, 1-based index, ub): DOCP_variables.jl:88-98
    double lb;
    int idx;
    double ub;
};

struct ProblemMeta {  // host side only
    std::vector<double> default_params;
    std::vector<std::string> param_names;
    std::vector<Box> state_box, control_box, variable_box;
    std::vector<double> path_lb, path_ub, bc_lb, bc_ub;
    std::vector<double> init_state, init_control, init_variable;
};

// Run-time copy of a generated element's compile-time description, so that the host-side model
// (layout, bounds, sparsity structure, edge tables) is ordinary non-templated C++.
struct SlotTable {
    int n = 0;
    std::vector<short> e;  // n x 3: (local row | a, local col | b | kv, slot)
    int operator()(int i, int k) const { return e[3 * i + k]; }
};

struct ElemInfo {
    std::string problem, scheme;
    int scheme_id = 0, pattern_mode = 0;
    int nx = 0, nu = 0, nv = 0, np = 0, nb = 0, npar = 0, stage = 0;
    bool free_t0 = false, free_tf = false, lagrange = false, mayer = false, is_max = false, irk = false,
         stagewise = false, final_control = false, synthetic = false;
    int ctrl_blk = 0, blk = 0, eqs = 0, cb = 0, nown = 0, nnext = 0, nz = 0;
    bool path_shift = false; // implicit Euler with path constraints: a step's path rows are those of the NEXT node
    int control_steps = 1;   // controls per time step (DirectShooting: src/direct_shooting.jl:55-71, midpoint.jl:20)
    int j_block = 0, h_block = 0, j_first = 0, h_first = 0, jv_stride = 0, fp_nloc = 0, bc_nloc = 0;
    int n_dropped = 0;
    int nh_vv_emit = 0;   // (v,v) Hessian terms emitted by a step (0: no grid-wide sum for them)
    SlotTable j_own, j_next, j_v, h_own, h_next, h_vv, fp_j, fp_h, bc_j, bc_h, g_own, g_next, g_v, bc_g;
    SlotTable s_step_j, s_step_h, s_fp_j, s_fp_h, s_bc_j, s_bc_h;   // detected structure (pattern mode 2)
    ProblemMeta meta;
};

#define CTD_FILL_TABLE(info, E, NAME, field)                                 \
    do {                                                                     \
        (info).field.n = E::N##NAME;                                         \
        (info).field.e.resize(3 * (size_t)E::N##NAME);                       \
        for (int _i = 0; _i < E::N##NAME; ++_i)                              \
            for (int _k = 0; _k < 3; ++_k) (info).field.e[3 * _i + _k] = (short)E::NAME(_i, _k); \
    } while (0)

template <class E>
inline ElemInfo make_elem_info() {
    ElemInfo f;
    f.problem = E::PROBLEM;
    f.scheme = E::SCHEME;
    f.scheme_id = E::SCHEME_ID;
    f.pattern_mode = E::PATTERN_MODE;
    f.nx = E::NX; f.nu = E::NU; f.nv = E::NV; f.np = E::NP; f.nb = E::NB; f.npar = E::NPAR; f.stage = E::STAGE;
    f.free_t0 = E::FREE_T0; f.free_tf = E::FREE_TF; f.lagrange = E::LAGRANGE; f.mayer = E::MAYER;
    f.is_max = E::IS_MAX; f.irk = E::IRK; f.stagewise = E::STAGEWISE; f.final_control = E::FINAL_CONTROL;
    f.synthetic = E::SYNTHETIC;
    f.ctrl_blk = E::CTRL_BLK; f.blk = E::BLK; f.eqs = E::EQS; f.cb = E::CB;
    f.control_steps = E::CONTROL_STEPS;
    f.path_shift = E::PATH_SHIFT;
    f.nown = E::NOWN; f.nnext = E::NNEXT; f.nz = E::NZ;
    f.j_block = E::J_BLOCK; f.h_block = E::H_BLOCK; f.j_first = E::J_FIRST; f.h_first = E::H_FIRST;
    f.jv_stride = E::JV_STRIDE; f.fp_nloc = E::FP_NLOC; f.bc_nloc = E::BC_NLOC; f.n_dropped = E::N_DROPPED;
    f.nh_vv_emit = E::NH_VV_EMIT;
    CTD_FILL_TABLE(f, E, J_OWN, j_own);
    CTD_FILL_TABLE(f, E, J_NEXT, j_next);
    CTD_FILL_TABLE(f, E, J_V, j_v);
    CTD_FILL_TABLE(f, E, H_OWN, h_own);
    CTD_FILL_TABLE(f, E, H_NEXT, h_next);
    CTD_FILL_TABLE(f, E, H_VV, h_vv);
    CTD_FILL_TABLE(f, E, S_STEP_J, s_step_j);
    CTD_FILL_TABLE(f, E, S_STEP_H, s_step_h);
    CTD_FILL_TABLE(f, E, S_FP_J, s_fp_j);
    CTD_FILL_TABLE(f, E, S_FP_H, s_fp_h);
    CTD_FILL_TABLE(f, E, S_BC_J, s_bc_j);
    CTD_FILL_TABLE(f, E, S_BC_H, s_bc_h);
    CTD_FILL_TABLE(f, E, FP_J, fp_j);
    CTD_FILL_TABLE(f, E, FP_H, fp_h);
    CTD_FILL_TABLE(f, E, BC_J, bc_j);
    CTD_FILL_TABLE(f, E, BC_H, bc_h);
    CTD_FILL_TABLE(f, E, G_OWN, g_own);
    CTD_FILL_TABLE(f, E, G_NEXT, g_next);
    CTD_FILL_TABLE(f, E, G_V, g_v);
    CTD_FILL_TABLE(f, E, BC_G, bc_g);
    E::meta(f.meta);
    return f;
}

}  // namespace ctd
