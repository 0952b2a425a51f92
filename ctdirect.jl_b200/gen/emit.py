"""CUDA C emission for the generated (OCP, scheme) element structs.

The emitted header defines one struct per pair with
  * compile-time dimensions, value-slot tables and problem metadata,
  * `step<WC,WJ,WH>(...)`: constraint rows (in the reference's association order), Jacobian and
    Lagrangian-Hessian values of one time step, streamed into a sink `S` by slot number,
  * `fpath`, `bc`: final path constraints and boundary constraints + Mayer cost (edge elements),
  * `cost_step`, `mayer`, `grad_step`, `grad_bc`: objective and its gradient.
Kernels (csrc/ctd_kernels.cuh) provide the sinks; nothing here knows about memory layout beyond the
slot numbers.
"""
from __future__ import annotations

import math

import sympy as sp
from sympy.printing.c import C99CodePrinter
from sympy.printing.precedence import PRECEDENCE

from . import elements as el
from . import patterns as pat
from .ocps import OCP, ZERO


class CudaPrinter(C99CodePrinter):
    def _print_Pow(self, expr):
        b, e = expr.base, expr.exp
        if e.is_Integer and e != 0 and abs(int(e)) <= 4:
            bs = self.parenthesize(b, PRECEDENCE["Mul"])
            prod = "*".join([bs] * abs(int(e)))
            if e > 0:
                return f"({prod})"
            return f"(1.0/({prod}))"
        if e == sp.Rational(1, 2):
            return f"sqrt({self._print(b)})"
        if e == sp.Rational(-1, 2):
            return f"(1.0/sqrt({self._print(b)}))"
        return super()._print_Pow(expr)

    def _print_rcp(self, expr):
        return f"(1.0/({self._print(expr.args[0])}))"

    def _print_Integer(self, expr):
        return f"{int(expr)}.0" if abs(int(expr)) < 2**53 else super()._print_Integer(expr)

    def _print_Float(self, expr):
        return repr(float(expr))


_PR = CudaPrinter()


def cc(e):
    return _PR.doprint(sp.sympify(e))


class rcp(sp.Function):
    """1/x as an opaque function, so that CSE shares ONE reciprocal per base between all negative powers
    of it (an FP64 division costs ~30 instructions on the GPU; x**-3 becomes rcp(x)**3)."""
    nargs = 1


def _share_reciprocals(e):
    e = sp.sympify(e)
    return e.replace(lambda q: q.is_Pow and q.exp.is_Integer and q.exp < 0,
                     lambda q: sp.Pow(rcp(q.base), -q.exp))


def cse_block(exprs, tmp_prefix):
    """Return (list of 'const double t = ...;' lines, list of C strings for the outputs)."""
    exprs = [_share_reciprocals(e) for e in exprs]
    if not exprs:
        return [], []
    repl, red = sp.cse(exprs, symbols=sp.numbered_symbols(tmp_prefix), order="none")
    lines = [f"const double {cc(s)} = {cc(e)};" for s, e in repl]
    return lines, [cc(e) for e in red]


def dlit(x: float) -> str:
    if x == float("inf"):
        return "CTD_INF"
    if x == float("-inf"):
        return "(-CTD_INF)"
    return repr(float(x))


def butcher_values(s: int):
    """Numerical tableau from the reference's own expressions (irk.jl:41-43,78-80,116-122)."""
    if s == 1:
        return [[0.5]], [1.0], [0.5]
    if s == 2:
        a = [[0.25, 0.25 - math.sqrt(3) / 6], [0.25 + math.sqrt(3) / 6, 0.25]]
        return a, [0.5, 0.5], [0.5 - math.sqrt(3) / 6, 0.5 + math.sqrt(3) / 6]
    a = [
        [5.0 / 36.0, 2 / 9 - math.sqrt(15) / 15, 5 / 36 - math.sqrt(15) / 30],
        [5.0 / 36.0 + math.sqrt(15) / 24, 2.0 / 9.0, 5.0 / 36.0 - math.sqrt(15) / 24],
        [5 / 36 + math.sqrt(15) / 30, 2 / 9 + math.sqrt(15) / 15, 5.0 / 36.0],
    ]
    return a, [5.0 / 18.0, 4.0 / 9.0, 5.0 / 18.0], [0.5 - 0.1 * math.sqrt(15), 0.5, 0.5 + 0.1 * math.sqrt(15)]


def _named(prefix, n):
    return [sp.Symbol(f"{prefix}{k}") for k in range(n)]


def _dyn_lines(ocp: OCP, tag, tname, xnames, unames, vsyms, psyms):
    """C lines computing {tag}f0.. = dynamics(t, x, u, v) with named inputs."""
    t = sp.Symbol(tname)
    f = [sp.sympify(e).subs(ZERO, 0) for e in
         ocp.dynamics(t, [sp.Symbol(n) for n in xnames], [sp.Symbol(n) for n in unames], vsyms, psyms)]
    lines, outs = cse_block(f, f"{tag}w")
    for k, o in enumerate(outs):
        lines.append(f"const double {tag}f{k} = {o};")
    return lines


def _cons_lines(ocp: OCP, E: el.StepElement):
    """Constraint rows in the reference's evaluation order.  The dynamics are evaluated at the
    intermediate points defined by the prologue (E.inter_code); CTD_ADD / CTD_SUB / CTD_MUL are
    non-contracting (no FMA), so the scheme formulas round exactly like the reference's Julia code."""
    L = E.L
    d = L.dims
    nx, nu, nv, npc, s = d.nx, d.nu, d.nv, d.np, L.stage
    z = [f"z{a}" for a in range(L.nz)]
    vs = E.z[L.n_own + L.n_next:]
    xi = z[0:nx]
    xn = z[L.n_own:L.n_own + nx]
    out = []
    if pat.is_midpoint(E.scheme) and L.control_steps == 1:
        # midpoint.jl:139  c = xip1 - (xi + hi*f(ts, xs, ui, v))
        out += _dyn_lines(ocp, "c_", "m_t", [f"m_x{k}" for k in range(nx)], z[nx:nx + nu], vs, E.p)
        for k in range(nx):
            out.append(f"s.template c<{k}>(CTD_SUB({xn[k]}, CTD_ADD({xi[k]}, CTD_MUL(hh, c_f{k}))));")
    elif pat.is_midpoint(E.scheme):
        # midpoint.jl:145-154  x_next = xi; x_next += hi * f(ts, xs, u_ij, v) for j = 1..control_steps; c = xip1 - x_next
        acc = list(xi)
        for j in range(L.control_steps):
            out += _dyn_lines(ocp, f"c{j}_", "m_t", [f"m_x{k}" for k in range(nx)], z[nx + j * nu:nx + (j + 1) * nu], vs, E.p)
            acc = [f"CTD_ADD({acc[k]}, CTD_MUL(m_h, c{j}_f{k}))" for k in range(nx)]
        for k in range(nx):
            out.append(f"s.template c<{k}>(CTD_SUB({xn[k]}, {acc[k]}));")
    elif E.scheme in ("euler", "euler_implicit"):
        # euler.jl:140-159  c = xip1 - (xi + hi*f), f at node i (explicit) or node i+1 (implicit)
        if E.scheme == "euler":
            out += _dyn_lines(ocp, "c_", "ti", xi, z[nx:nx + nu], vs, E.p)
        else:
            out += _dyn_lines(ocp, "c_", "tp", xn, z[nx:nx + nu], vs, E.p)
        for k in range(nx):
            out.append(f"s.template c<{k}>(CTD_SUB({xn[k]}, CTD_ADD({xi[k]}, CTD_MUL(hh, c_f{k}))));")
    elif E.scheme == "trapeze":
        # trapeze.jl:128-140  c = xip1 - (xi + half_hi*(f_i + f_ip1))
        out += _dyn_lines(ocp, "ci_", "ti", xi, z[nx:nx + nu], vs, E.p)
        out += _dyn_lines(ocp, "cn_", "tp", xn, z[L.n_own + nx:L.n_own + nx + nu], vs, E.p)
        out.append("const double c_hh2 = CTD_MUL(0.5, CTD_SUB(tp, ti));")
        for k in range(nx):
            out.append(f"s.template c<{k}>(CTD_SUB({xn[k]}, CTD_ADD({xi[k]}, CTD_MUL(c_hh2, CTD_ADD(ci_f{k}, cn_f{k})))));")
    else:
        cb = L.control_block
        kname = lambda j, k: z[nx + cb + j * nx + k]
        for j in range(s):
            # irk.jl:286-302 / irk_stagewise.jl:440-454  c_stage = k_j - f(tij, xij, u, v)
            un = z[nx + j * nu:nx + (j + 1) * nu] if L.stagewise else z[nx:nx + nu]
            out += _dyn_lines(ocp, f"c{j}_", f"s{j}_t", [f"s{j}_x{k}" for k in range(nx)], un, vs, E.p)
            for k in range(nx):
                out.append(f"s.template c<{nx + j * nx + k}>(CTD_SUB({kname(j, k)}, c{j}_f{k}));")
        for k in range(nx):
            e = f"CTD_MUL(BB0, {kname(0, k)})"
            for j in range(1, s):
                e = f"CTD_ADD({e}, CTD_MUL(BB{j}, {kname(j, k)}))"
            out.append(f"s.template c<{k}>(CTD_SUB({xn[k]}, CTD_ADD({xi[k]}, CTD_MUL(hh, {e}))));")
    if npc:
        if E.path_shift:   # rows of the NEXT node: g(t_{i+1}, x_{i+1}, U_i, v)  (euler.jl:66-69)
            g = ocp.path(sp.Symbol("tp"), [sp.Symbol(n) for n in xn], list(E.u_node), vs, E.p)
        else:
            g = ocp.path(sp.Symbol("ti"), [sp.Symbol(n) for n in xi], list(E.u_node), vs, E.p)
        lines, outs = cse_block(g, "c_pw")
        out += lines
        for r, o in enumerate(outs):
            out.append(f"s.template c<{L.eqs_block + r}>({o});")
    return out


def _prologue(nz, nlam, npar, ts: el.TimeSyms, use_time=True, inter_code=()):
    out = [f"const double z{a} = z[{a}];" for a in range(nz)]
    out += [f"const double lam{r} = lam[{r}];" for r in range(nlam)]
    out += [f"const double p{k} = p[{k}];" for k in range(npar)]
    if use_time:
        out += ["const double ti = T.ti, tp = T.tp, hh = T.hh;"]
        for k in sorted(ts.dti):
            out.append(f"const double dti{k} = T.dti[{k}], dtp{k} = T.dtp[{k}], dh{k} = T.dh[{k}];")
    out += list(inter_code)
    return out


def _table(name, entries, cols=3):
    """Value-slot table `NAME(n, k)`: a constexpr accessor (usable from host and device code; a static
    constexpr array member would be an odr-use in device code, which nvcc rejects)."""
    n = len(entries)
    rows = [e[:cols] for e in entries] if n else [(0,) * cols]
    body = ", ".join("{" + ", ".join(str(int(x)) for x in e) + "}" for e in rows)
    return (f"  static constexpr int N{name} = {n};\n"
            f"  __host__ __device__ static constexpr int {name}(int n, int k) {{\n"
            f"    constexpr short t[{max(n, 1)}][{cols}] = {{{body}}};\n    return t[n][k];\n  }}\n")


def _indent(lines, n=4):
    pad = " " * n
    return "".join(pad + l + "\n" for l in lines)


def struct_name(ocp_name, scheme):
    return f"E_{ocp_name}__{scheme}"


def emit_pair(ocp: OCP, scheme: str, pattern_mode: str = "reference") -> str:
    E = el.build_step_element(ocp, scheme, pattern_mode)
    FP = el.build_final_path(ocp, scheme, E)
    BC = el.build_boundary(ocp, E)
    L = E.L
    d = L.dims
    s = L.stage
    name = struct_name(ocp.name, scheme) + ("" if pattern_mode == "reference" else "__" + pattern_mode)
    scheme_id = list(pat.SCHEMES).index(scheme)
    o = []
    o.append(f"// AUTO-GENERATED by ctdirect.jl_b200/gen/generate.py for problem '{ocp.name}', scheme '{scheme}',\n"
             f"// pattern_mode '{pattern_mode}'.  Do not edit by hand.\n#pragma once\n#include \"../ctd_elem_common.cuh\"\n")
    o.append("namespace ctd_gen {\n")
    o.append(f"struct {name} {{\n")
    o.append(f"  static constexpr const char* PROBLEM = \"{ocp.name}\";\n  static constexpr const char* SCHEME = \"{scheme}\";\n")
    o.append(f"  static constexpr int SCHEME_ID = {scheme_id}, PATTERN_MODE = {['reference', 'corrected', 'optimized'].index(pattern_mode)};\n")
    o.append(f"  static constexpr int NX = {d.nx}, NU = {d.nu}, NV = {d.nv}, NP = {d.np}, NB = {d.nb}, NPAR = {ocp.npar}, STAGE = {s}, CONTROL_STEPS = {L.control_steps};\n")
    b = lambda x: "true" if x else "false"
    o.append(f"  static constexpr bool FREE_T0 = {b(ocp.free_t0)}, FREE_TF = {b(ocp.free_tf)}, LAGRANGE = {b(d.lagrange)}, "
             f"MAYER = {b(d.mayer)}, IS_MAX = {b(ocp.criterion == 'max')}, IRK = {b(L.irk)}, STAGEWISE = {b(L.stagewise)}, "
             f"FINAL_CONTROL = {b(L.final_control)}, SYNTHETIC = {b(ocp.synthetic)}, "
             f"EULER_IMPLICIT = {b(scheme == 'euler_implicit')}, PATH_SHIFT = {b(E.path_shift)};\n")
    tk = lambda spec: (0 if spec[0] == "const" else 1 if spec[0] == "var" else 2)
    o.append(f"  static constexpr int T0_KIND = {tk(ocp.t0)}, TF_KIND = {tk(ocp.tf)};  // 0 const, 1 v[idx], 2 p[idx]\n")
    o.append(f"  static constexpr int T0_IDX = {int(ocp.t0[1]) if ocp.t0[0] != 'const' else -1}, TF_IDX = {int(ocp.tf[1]) if ocp.tf[0] != 'const' else -1};\n")
    o.append(f"  static constexpr double T0_CONST = {dlit(ocp.t0[1]) if ocp.t0[0] == 'const' else '0.0'}, TF_CONST = {dlit(ocp.tf[1]) if ocp.tf[0] == 'const' else '0.0'};\n")
    o.append(f"  static constexpr int CTRL_BLK = {L.control_block}, BLK = {L.step_block}, EQS = {L.eqs_block}, CB = {L.cblock};\n")
    o.append(f"  static constexpr int NOWN = {L.n_own}, NNEXT = {L.n_next}, NZ = {L.nz};\n")
    o.append(f"  static constexpr int J_BLOCK = {E.j_block}, H_BLOCK = {E.h_block}, J_FIRST = {E.j_first}, H_FIRST = {E.h_first}, JV_STRIDE = {E.jv_stride};\n")
    o.append(f"  static constexpr int NH_VV_SLOTS = {d.nv * (d.nv + 1) // 2};\n")
    # (v,v) Hessian terms / d cost / d v terms that a STEP emits: when there are none (autonomous dynamics with the
    # free time entering only through h, Mayer cost) the kernels drop the grid-wide reduction of those slots
    o.append(f"  static constexpr int NH_VV_EMIT = {len(E.h_vv)};\n")
    o.append(f"  static constexpr int FP_NLOC = {FP.nloc}, BC_NLOC = {BC.nloc};\n")
    o.append(f"  static constexpr int N_DROPPED = {len(E.dropped)};  // mathematically nonzero entries absent from the pattern\n")
    # tables: {local row / a, local col / b, slot}
    o.append(_table("J_OWN", E.j_own))
    o.append(_table("J_NEXT", E.j_next))
    o.append(_table("J_V", [(r, a - L.n_own - L.n_next, off) for (r, a, off, _) in E.j_v]))
    o.append(_table("H_OWN", E.h_own))
    o.append(_table("H_NEXT", E.h_next))
    o.append(_table("H_VV", E.h_vv_table))
    # detected structure (":optimized" pattern mode only): local (row, variable) / (variable, variable) lists
    S = E.structure or {}
    for nm, key in (("S_STEP_J", "step_J"), ("S_STEP_H", "step_H"), ("S_FP_J", "fp_J"), ("S_FP_H", "fp_H"),
                    ("S_BC_J", "bc_J"), ("S_BC_H", "bc_H")):
        o.append(_table(nm, [(a, b, 0) for (a, b) in S.get(key, [])]))
    o.append(_table("FP_J", FP.jac))
    o.append(_table("FP_H", FP.hess))
    o.append(_table("BC_J", BC.jac))
    o.append(_table("BC_H", BC.hess))
    # Butcher
    if s:
        A, Bw, Cw = butcher_values(s)
        consts = []
        for j in range(s):
            for l in range(s):
                consts.append(f"BA{j}{l} = {A[j][l].hex()}")
            consts.append(f"BB{j} = {float(Bw[j]).hex()}")
            consts.append(f"BC{j} = {float(Cw[j]).hex()}")
        o.append("  static constexpr double " + ", ".join(consts) + ";\n")
    # bounds of the constraint rows of one step (0 for the state / stage equations, the path-constraint
    # bounds after them) and of the boundary rows: DOCP_functions.jl:163-191
    def _dtab(fname, vals):
        vals = list(vals) or [0.0]
        return (f"  __host__ __device__ static constexpr double {fname}(int r) {{\n"
                f"    constexpr double t[{len(vals)}] = {{{', '.join(dlit(v) for v in vals)}}};\n    return t[r];\n  }}\n")
    o.append(_dtab("CON_LB", [0.0] * L.eqs_block + list(ocp.path_lb)))
    o.append(_dtab("CON_UB", [0.0] * L.eqs_block + list(ocp.path_ub)))
    o.append(_dtab("BC_LB", ocp.bc_lb))
    o.append(_dtab("BC_UB", ocp.bc_ub))
    bw = [float(x).hex() for x in (butcher_values(s)[1] if s else [])] + ["0.0"] * (3 - s)
    o.append("  __host__ __device__ static constexpr double BB(int j) {\n"
             f"    constexpr double t[3] = {{{', '.join(bw)}}};\n    return t[j];\n  }}\n")
    # host-side metadata
    o.append("  static void meta(ctd::ProblemMeta& m) {\n")
    o.append("    m.default_params = {" + ", ".join(dlit(v) for _, v in ocp.params) + "};\n")
    o.append("    m.param_names = {" + ", ".join(f'"{n}"' for n, _ in ocp.params) + "};\n")
    for nm, box in (("state_box", ocp.state_box), ("control_box", ocp.control_box), ("variable_box", ocp.variable_box)):
        o.append(f"    m.{nm} = {{" + ", ".join(f"{{{dlit(lb)}, {idx}, {dlit(ub)}}}" for lb, idx, ub in box) + "};\n")
    for nm, vec in (("path_lb", ocp.path_lb), ("path_ub", ocp.path_ub), ("bc_lb", ocp.bc_lb), ("bc_ub", ocp.bc_ub),
                    ("init_state", ocp.init_state or []), ("init_control", ocp.init_control or []),
                    ("init_variable", ocp.init_variable or [])):
        o.append(f"    m.{nm} = {{" + ", ".join(dlit(v) for v in vec) + "};\n")
    o.append("  }\n\n")

    # ------------------------------------------------------------------ step()
    o.append("  template <bool WC, bool WJ, bool WH, class S>\n"
             "  __device__ __forceinline__ static void step(const double* __restrict__ z, const ctd::TimeArgs<NV>& T,\n"
             "      const double* __restrict__ p, const double* __restrict__ lam, const double sg, S& s) {\n")
    o.append(_indent(_prologue(L.nz, L.cblock, ocp.npar, E.ts, inter_code=E.inter_code)))
    o.append("    if constexpr (WC) {\n" + _indent(_cons_lines(ocp, E), 6) + "    }\n")
    jex = [e[3] for e in E.j_own] + [e[3] for e in E.j_next] + [e[3] for e in E.j_v]
    hex_ = [e[3] for e in E.h_own] + [e[3] for e in E.h_next] + [e[3] for e in E.h_vv]
    lines, outs = cse_block(jex + hex_, "w")
    o.append("    if constexpr (WJ || WH) {\n")
    o.append(_indent(lines, 6))
    k = 0
    o.append("      if constexpr (WJ) {\n")
    for tag, lst in (("jo", E.j_own), ("jn", E.j_next), ("jv", E.j_v)):
        for n, e in enumerate(lst):
            o.append(f"        s.template {tag}<{n}>({outs[k]});\n")
            k += 1
    o.append("      }\n      if constexpr (WH) {\n")
    for tag, lst in (("ho", E.h_own), ("hn", E.h_next)):
        for n, e in enumerate(lst):
            o.append(f"        s.template {tag}<{n}>({outs[k]});\n")
            k += 1
    for e in E.h_vv:
        o.append(f"        s.template hv<{e[2]}>({outs[k]});\n")
        k += 1
    o.append("      }\n    }\n  }\n\n")

    # ------------------------------------------------------------------ fpath()
    o.append("  template <bool WC, bool WJ, bool WH, class S>\n"
             "  __device__ __forceinline__ static void fpath(const double* __restrict__ z, const ctd::TimeArgs<NV>& T,\n"
             "      const double* __restrict__ p, const double* __restrict__ lam, S& s) {\n")
    if d.np:
        o.append(_indent(_prologue(FP.nloc, d.np, ocp.npar, E.ts)))
        lines, outs = cse_block(FP.rows + [e[3] for e in FP.jac] + [e[3] for e in FP.hess], "w")
        o.append(_indent(lines))
        k = 0
        o.append("    if constexpr (WC) {\n")
        for r in range(FP.nrows):
            o.append(f"      s.template c<{r}>({outs[k]});\n")
            k += 1
        o.append("    }\n    if constexpr (WJ) {\n")
        for n, e in enumerate(FP.jac):
            o.append(f"      s.template j<{n}>({outs[k]});\n")
            k += 1
        o.append("    }\n    if constexpr (WH) {\n")
        for n, e in enumerate(FP.hess):
            o.append(f"      s.template h<{n}>({outs[k]});\n")
            k += 1
        o.append("    }\n")
    o.append("  }\n\n")

    # ------------------------------------------------------------------ bc()
    o.append("  template <bool WC, bool WJ, bool WH, class S>\n"
             "  __device__ __forceinline__ static void bc(const double* __restrict__ z, const double* __restrict__ p,\n"
             "      const double* __restrict__ lam, const double sg, S& s) {\n")
    o.append(_indent(_prologue(BC.nloc, d.nb, ocp.npar, E.ts, use_time=False)))
    lines, outs = cse_block(BC.rows + [e[3] for e in BC.jac] + [e[3] for e in BC.hess], "w")
    o.append(_indent(lines))
    k = 0
    o.append("    if constexpr (WC) {\n")
    for r in range(BC.nrows):
        o.append(f"      s.template c<{r}>({outs[k]});\n")
        k += 1
    o.append("    }\n    if constexpr (WJ) {\n")
    for n, e in enumerate(BC.jac):
        o.append(f"      s.template j<{n}>({outs[k]});\n")
        k += 1
    o.append("    }\n    if constexpr (WH) {\n")
    for n, e in enumerate(BC.hess):
        o.append(f"      s.template h<{n}>({outs[k]});\n")
        k += 1
    o.append("    }\n  }\n\n")

    # ------------------------------------------------------------------ objective pieces
    o.append("  __device__ __forceinline__ static double cost_step(const double* __restrict__ z, const ctd::TimeArgs<NV>& T,\n"
             "      const double* __restrict__ p) {\n")
    if d.lagrange:
        o.append(_indent(_prologue(L.nz, 0, ocp.npar, E.ts, inter_code=E.inter_code)))
        lines, outs = cse_block([E.cost], "w")
        o.append(_indent(lines))
        o.append(f"    return {outs[0]};\n")
    else:
        o.append("    return 0.0;\n")
    o.append("  }\n")
    o.append("  __device__ __forceinline__ static double mayer(const double* __restrict__ z, const double* __restrict__ p) {\n")
    if d.mayer:
        o.append(_indent(_prologue(BC.nloc, 0, ocp.npar, E.ts, use_time=False)))
        lines, outs = cse_block([BC.cost], "w")
        o.append(_indent(lines))
        o.append(f"    return {outs[0]};\n")
    else:
        o.append("    return 0.0;\n")
    o.append("  }\n")
    o.append(_table("G_OWN", E.g_own))
    o.append(_table("G_NEXT", E.g_next))
    o.append(_table("G_V", E.g_v))
    o.append(_table("BC_G", BC.grad))
    o.append("  template <class S>\n  __device__ __forceinline__ static void grad_step(const double* __restrict__ z, const ctd::TimeArgs<NV>& T,\n"
             "      const double* __restrict__ p, S& s) {\n")
    if d.lagrange:
        o.append(_indent(_prologue(L.nz, 0, ocp.npar, E.ts, inter_code=E.inter_code)))
        lines, outs = cse_block([e[3] for e in E.g_own] + [e[3] for e in E.g_next] + [e[3] for e in E.g_v], "w")
        o.append(_indent(lines))
        k = 0
        for tag, lst in (("go", E.g_own), ("gn", E.g_next), ("gv", E.g_v)):
            for n, e in enumerate(lst):
                o.append(f"    s.template {tag}<{n}>({outs[k]});\n")
                k += 1
    o.append("  }\n")
    o.append("  template <class S>\n  __device__ __forceinline__ static void grad_bc(const double* __restrict__ z, const double* __restrict__ p, S& s) {\n")
    if d.mayer and BC.grad:
        o.append(_indent(_prologue(BC.nloc, 0, ocp.npar, E.ts, use_time=False)))
        lines, outs = cse_block([e[3] for e in BC.grad], "w")
        o.append(_indent(lines))
        for n, e in enumerate(BC.grad):
            o.append(f"    s.template g<{n}>({outs[n]});\n")
    o.append("  }\n")
    o.append("};\n}  // namespace ctd_gen\n")
    return "".join(o), name, E
