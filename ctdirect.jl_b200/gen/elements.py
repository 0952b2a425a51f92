"""Symbolic per-step "elements" of a discretised OCP and their value-slot tables.

For one (OCP, scheme) pair this module builds, with sympy,
  * the constraint rows of one time step (state equation, stage equations, path constraints) --
    stepStateConstraints! / stepPathConstraints! of the reference (src/ode/midpoint.jl:124-156,
    trapeze.jl:118-142, irk.jl:236-308, irk_stagewise.jl:394-460, DOCP_functions.jl:122-140),
  * the running-cost term of the step (integral(): midpoint.jl:79-116, trapeze.jl:78-110,
    irk.jl:179-228, irk_stagewise.jl:344-384),
  * their analytic first and second derivatives with respect to the step's local variables
    z = (x_i, u_i[, k_i], x_{i+1}[, u_{i+1}], v),
and, by probing the reference-faithful patterns (gen/patterns.py) on a tiny grid, the offset of
every derivative value inside the CSC value arrays of an interior column block.

Time enters through t_i, t_{i+1} and h_i = t_{i+1} - t_i, which the kernels compute exactly like
the reference (t = t0 + s*(tf - t0), DOCP_data.jl:456); they are kept as opaque symbols whose
derivatives with respect to free t0 / tf are supplied as seeds (s_i, 1 - s_i, and their
differences), mirroring what forward-mode AD of the reference code produces.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Tuple

import sympy as sp

from . import patterns as pat
from . import trace as tr
from .ocps import OCP, ZERO

PROBE_N = 6
PROBE_I = 3


def butcher_symbols(s: int):
    """Opaque symbols for the tableau (values are defined in the emitted code from the reference's
    own expressions, irk.jl:41-43,78-80,116-122)."""
    a = [[sp.Symbol(f"BA{j}{l}") for l in range(s)] for j in range(s)]
    b = [sp.Symbol(f"BB{j}") for j in range(s)]
    c = [sp.Symbol(f"BC{j}") for j in range(s)]
    return a, b, c


@dataclass
class TimeSyms:
    ti: sp.Symbol
    tp: sp.Symbol
    h: sp.Symbol
    # seeds: d ti / d v_k etc.; zero when v_k is not a free time
    dti: Dict[int, sp.Symbol]
    dtp: Dict[int, sp.Symbol]
    dh: Dict[int, sp.Symbol]


class Differ:
    """Total derivatives with respect to local variables.  Time symbols and *intermediate* symbols
    (midpoint / stage states and times, which the kernels evaluate once, in the reference's association
    order, and feed to the nonlinear OCP functions) are folded in by the chain rule."""

    def __init__(self, z, nv_off, nv, ts: TimeSyms, inter=None):
        self.z = z
        self.nv_off = nv_off
        self.nv = nv
        self.ts = ts
        self.inter = list(inter or [])   # [(symbol, defining expression in z / time / tableau symbols)]

    def _base(self, e, a):
        e = sp.sympify(e)
        r = sp.diff(e, self.z[a])
        k = a - self.nv_off
        if 0 <= k < self.nv:
            if k in self.ts.dti:
                r = r + sp.diff(e, self.ts.ti) * self.ts.dti[k]
            if k in self.ts.dtp:
                r = r + sp.diff(e, self.ts.tp) * self.ts.dtp[k]
            if k in self.ts.dh:
                r = r + sp.diff(e, self.ts.h) * self.ts.dh[k]
        return r

    def d(self, e, a):
        e = sp.sympify(e)
        r = self._base(e, a)
        for sym, expr in self.inter:
            de = sp.diff(e, sym)
            if de != 0:
                r = r + de * self._base(expr, a)
        return r


@dataclass
class StepElement:
    ocp: OCP
    scheme: str
    L: pat.Layout                      # probe layout (dims + block sizes)
    pattern_mode: str
    z: List[sp.Symbol] = field(default_factory=list)
    lam: List[sp.Symbol] = field(default_factory=list)
    p: List[sp.Symbol] = field(default_factory=list)
    sg: sp.Symbol = None
    ts: TimeSyms = None
    rows: List[sp.Expr] = field(default_factory=list)        # constraint rows (for derivatives)
    cost: sp.Expr = sp.Integer(0)                            # running cost of the step
    # value tables: lists of (local row / a, local col / b, offset, expr)
    j_own: list = field(default_factory=list)
    j_next: list = field(default_factory=list)
    j_v: list = field(default_factory=list)
    h_own: list = field(default_factory=list)
    h_next: list = field(default_factory=list)
    h_vv: list = field(default_factory=list)
    g_own: list = field(default_factory=list)
    g_next: list = field(default_factory=list)
    g_v: list = field(default_factory=list)
    j_block: int = 0
    h_block: int = 0
    j_first: int = 0      # base offset of column block 2 (= size of column block 1)
    h_first: int = 0
    jv_stride: int = 0    # values per step inside one v column
    dropped: list = field(default_factory=list)
    # c-value code pieces (dynamics evaluations kept apart so that the scheme formula can be
    # emitted in the reference's association order)
    cons_code: dict = field(default_factory=dict)
    inter: list = field(default_factory=list)        # [(symbol, expr)] intermediates (see Differ)
    inter_code: list = field(default_factory=list)   # C lines defining them in the reference's order
    u_node: list = field(default_factory=list)       # control used by the path constraints of the step
    path_shift: bool = False    # implicit Euler: the element's path rows are those of the NEXT node (see build_step_element)
    rows_traced: list = field(default_factory=list)  # rows / cost before structural zeros are dropped
    cost_traced: sp.Expr = sp.Integer(0)
    structure: dict = None                           # detected structure (pattern_mode 'optimized')
    h_vv_table: list = field(default_factory=list)   # (kv1, kv2, slot) of the (v,v) Hessian entries in the pattern


def _time_syms(ocp: OCP, prefix=""):
    ts = TimeSyms(sp.Symbol(prefix + "ti"), sp.Symbol(prefix + "tp"), sp.Symbol(prefix + "hh"), {}, {}, {})
    for which, spec in ((0, ocp.t0), (1, ocp.tf)):
        if spec[0] == "var":
            k = spec[1]
            ts.dti[k] = sp.Symbol(f"{prefix}dti{k}")
            ts.dtp[k] = sp.Symbol(f"{prefix}dtp{k}")
            ts.dh[k] = sp.Symbol(f"{prefix}dh{k}")
    return ts


def dims_of(ocp: OCP) -> pat.Dims:
    return pat.Dims(ocp.nx, ocp.nu, ocp.nv, ocp.np, ocp.nb, ocp.lagrange is not None, ocp.mayer is not None)


def build_step_element(ocp: OCP, scheme: str, pattern_mode: str = "reference") -> StepElement:
    d = dims_of(ocp)
    L = pat.Layout(scheme, d, PROBE_N)
    nx, nu, nv, npc = d.nx, d.nu, d.nv, d.np
    s = L.stage
    E = StepElement(ocp, scheme, L, pattern_mode)
    E.z = [sp.Symbol(f"z{a}") for a in range(L.nz)]
    E.lam = [sp.Symbol(f"lam{r}") for r in range(L.cblock)]
    E.p = [sp.Symbol(f"p{k}") for k in range(ocp.npar)]
    E.sg = sp.Symbol("sg")
    E.ts = _time_syms(ocp)
    z, p, ts = E.z, E.p, E.ts
    v = z[L.n_own + L.n_next:]
    xi = z[0:nx]
    xn = z[L.n_own:L.n_own + nx]
    ti, tp, h = ts.ti, ts.tp, ts.h
    half = sp.Rational(1, 2)

    def dyn(t, x, u):
        return [sp.sympify(e) for e in ocp.dynamics(t, x, u, v, p)]

    def lag(t, x, u):
        return sp.sympify(ocp.lagrange(t, x, u, v, p)) if ocp.lagrange else sp.Integer(0)

    rows = []
    cost = sp.Integer(0)
    inter, icode = [], []
    zn = [f"z{a}" for a in range(L.nz)]
    if pat.is_midpoint(scheme):
        # midpoint.jl:54-66: ts = 0.5*(ti + tip1), xs = 0.5*(xi + xip1)
        ncs = L.control_steps
        us = [z[nx + j * nu:nx + (j + 1) * nu] for j in range(ncs)]
        tsm = sp.Symbol("m_t")
        inter.append((tsm, half * (ti + tp)))
        icode.append("const double m_t = CTD_MUL(0.5, CTD_ADD(ti, tp));")
        xs = []
        for k in range(nx):
            sk = sp.Symbol(f"m_x{k}")
            xs.append(sk)
            inter.append((sk, half * (xi[k] + xn[k])))
            icode.append(f"const double m_x{k} = CTD_MUL(0.5, CTD_ADD({zn[k]}, {zn[L.n_own + k]}));")
        if ncs == 1:
            f = dyn(tsm, xs, us[0])
            rows = [xn[k] - (xi[k] + h * f[k]) for k in range(nx)]
            cost = h * lag(tsm, xs, us[0])
        else:
            # several controls per step (midpoint.jl:57-65,98-113,145-154): hi = (tip1 - ti) / control_steps; EVERY
            # sub-step evaluates the dynamics at the midpoint (ts, xs) of the whole step, with its own control, and the
            # state is advanced sub-step by sub-step; the running cost uses the sub-step midpoints tij = ti + (j - 0.5) hi
            hs = sp.Symbol("m_h")
            inter.append((hs, h / ncs))
            icode.append(f"const double m_h = hh / {float(ncs)!r};")
            rows = list(xi)
            for j in range(ncs):
                f = dyn(tsm, xs, us[j])
                rows = [rows[k] + hs * f[k] for k in range(nx)]
            rows = [xn[k] - rows[k] for k in range(nx)]
            cost = sp.Integer(0)
            if ocp.lagrange:
                for j in range(ncs):
                    tij = sp.Symbol(f"m_t{j}")
                    inter.append((tij, ti + (j + half) * h / ncs))   # in terms of h: intermediates do not nest
                    icode.append(f"const double m_t{j} = CTD_ADD(ti, CTD_MUL({j + 0.5!r}, m_h));")
                    cost = cost + hs * lag(tij, xs, us[j])
        u_node = us[0]       # path constraints read the first control of the step (common.jl:140-155, j = 1)
    elif scheme in ("euler", "euler_implicit"):
        # euler.jl:79-159: f at (t_i, x_i, U_i) (explicit) or at (t_{i+1}, x_{i+1}, U_i) (implicit)
        ui = z[nx:nx + nu]
        if scheme == "euler":
            f = dyn(ti, xi, ui)
            cost = h * lag(ti, xi, ui)
        else:
            if npc and pattern_mode == "reference":  # (corrected and optimized hold the dependency)
                raise NotImplementedError("implicit Euler with path constraints: node i reads U_{i-1} (euler.jl:66-69), "
                                          "a dependency the reference's :manual pattern does not hold; use 'corrected'")
            f = dyn(tp, xn, ui)
            cost = h * lag(tp, xn, ui)
        rows = [xn[k] - (xi[k] + h * f[k]) for k in range(nx)]
        u_node = ui
    elif scheme == "trapeze":
        ui = z[nx:nx + nu]
        un = z[L.n_own + nx:L.n_own + nx + nu]
        fi = dyn(ti, xi, ui)
        fn = dyn(tp, xn, un)
        rows = [xn[k] - (xi[k] + half * h * (fi[k] + fn[k])) for k in range(nx)]
        cost = half * h * (lag(ti, xi, ui) + lag(tp, xn, un))
        u_node = ui
    else:
        A, Bw, Cw = butcher_symbols(s)
        cb = L.control_block
        ks = [z[nx + cb + j * nx: nx + cb + (j + 1) * nx] for j in range(s)]
        kn = [[zn[nx + cb + j * nx + k] for k in range(nx)] for j in range(s)]
        if L.stagewise:
            us = [z[nx + j * nu: nx + (j + 1) * nu] for j in range(s)]
            # b-weighted average control (irk_stagewise.jl:197-205), accumulated left to right
            u_node = []
            for k in range(nu):
                sk = sp.Symbol(f"s_un{k}")
                inter.append((sk, sum(Bw[j] * us[j][k] for j in range(s))))
                e = f"CTD_MUL(BB0, {zn[nx + k]})"
                for j in range(1, s):
                    e = f"CTD_ADD({e}, CTD_MUL(BB{j}, {zn[nx + j * nu + k]}))"
                icode.append(f"const double s_un{k} = {e};")
                u_node.append(sk)
        else:
            us = [z[nx:nx + nu]] * s
            u_node = z[nx:nx + nu]
        stage_rows = []
        lsum = sp.Integer(0)
        for j in range(s):
            # irk.jl:262-302 / irk_stagewise.jl:424-454: tij = ti + c_j*hi ; xij = xi + sum_l (hi*a_jl)*k_l
            tij = sp.Symbol(f"s{j}_t")
            inter.append((tij, ti + Cw[j] * h))
            icode.append(f"const double s{j}_t = CTD_ADD(ti, CTD_MUL(BC{j}, hh));")
            xij = []
            for k in range(nx):
                sk = sp.Symbol(f"s{j}_x{k}")
                xij.append(sk)
                inter.append((sk, xi[k] + sum(h * A[j][l] * ks[l][k] for l in range(s))))
                e = zn[k]
                for l in range(s):
                    e = f"CTD_ADD({e}, CTD_MUL(CTD_MUL(hh, BA{j}{l}), {kn[l][k]}))"
                icode.append(f"const double s{j}_x{k} = {e};")
            f = dyn(tij, xij, us[j])
            stage_rows += [ks[j][k] - f[k] for k in range(nx)]
            lsum = lsum + Bw[j] * lag(tij, xij, us[j])
        state_rows = [xn[k] - (xi[k] + h * sum(Bw[j] * ks[j][k] for j in range(s))) for k in range(nx)]
        rows = state_rows + stage_rows
        cost = h * lsum
    E.inter, E.inter_code, E.u_node = inter, icode, list(u_node)
    # Implicit Euler: node i+1 reads the control of block i (euler.jl:66-69), i.e. the path constraints of node i+1
    # depend on (x_{i+1}, U_i) -- the inputs of step i.  The element of step i therefore carries the path rows of the
    # NEXT node: local rows EQS.. are global rows of the next step's block ("shifted" rows).  The path rows of node 1
    # (x_1, U_1) are evaluated by the edge CTA with the final-path code, those of node N+1 by the final-path element.
    E.path_shift = bool(npc) and scheme == "euler_implicit"
    if E.path_shift:
        rows += [sp.sympify(e) for e in ocp.path(tp, xn, u_node, v, p)]
    elif npc:
        rows += [sp.sympify(e) for e in ocp.path(ti, xi, u_node, v, p)]
    assert len(rows) == L.cblock
    E.rows_traced = [sp.sympify(r) for r in rows]
    E.cost_traced = sp.sympify(cost)
    rows = [sp.sympify(r).subs(ZERO, 0) for r in rows]
    cost = sp.sympify(cost).subs(ZERO, 0)
    E.rows = rows
    E.cost = sp.sympify(cost)

    # ------------------------------------------------------------------ value slots from the probe
    if pattern_mode == "optimized":
        E.structure = detect_structure(ocp, scheme, L, E)
        (jr, jc, jptr), (hr, hc, hptr) = optimized_patterns_for(ocp, scheme, L, E.structure)
    else:
        jr, jc, jptr = pat.jacobian_pattern(L, pattern_mode)
        hr, hc, hptr = pat.hessian_pattern(L, pattern_mode)
    B, CB = L.step_block, L.cblock
    nvar = L.nvar

    def jbase(i):  # 0-based position of the first value of column block i
        return int(jptr[(i - 1) * B])

    def hbase(i):
        return int(hptr[(i - 1) * B])

    E.j_block = jbase(PROBE_I + 1) - jbase(PROBE_I)
    E.h_block = hbase(PROBE_I + 1) - hbase(PROBE_I)
    E.j_first = jbase(2)
    E.h_first = hbase(2)
    for i in range(2, PROBE_N - 1):
        assert jbase(i + 1) - jbase(i) == E.j_block and hbase(i + 1) - hbase(i) == E.h_block

    def gcol(i, a):  # global 1-based column of local variable a at step i
        if a < L.n_own + L.n_next:
            return (i - 1) * B + a + 1
        return nvar - nv + (a - L.n_own - L.n_next) + 1

    def which(a):
        return "own" if a < L.n_own else ("next" if a < L.n_own + L.n_next else "v")

    D = Differ(z, L.n_own + L.n_next, nv, ts, inter)
    jv_stride = None
    if nv > 0:
        c0 = int(jptr[nvar - nv])
        # values per step inside a v column (CB for midpoint/IRK; NP for reference-mode trapeze)
        cnt = [int(((jr[c0:jptr[nvar - nv + 1]] > (i - 1) * CB) & (jr[c0:jptr[nvar - nv + 1]] <= i * CB)).sum())
               for i in (PROBE_I, PROBE_I + 1)]
        assert cnt[0] == cnt[1]
        jv_stride = cnt[0]
    E.jv_stride = jv_stride or 0

    def grow(i, r):  # 1-based global row of local row r of step i
        return (i - 1 + (1 if (E.path_shift and r >= L.eqs_block) else 0)) * CB + r + 1

    for r in range(CB):
        for a in range(L.nz):
            expr = sp.simplify(D.d(rows[r], a)) if False else D.d(rows[r], a)
            offs = []
            for i in (PROBE_I, PROBE_I + 1):
                pos = pat.find_pos(jr, jptr, grow(i, r), gcol(i, a))
                if pos < 0:
                    offs.append(None)
                    continue
                w = which(a)
                if w == "own":
                    offs.append(pos - jbase(i))
                elif w == "next":
                    offs.append(pos - jbase(i + 1))
                else:
                    kv = a - L.n_own - L.n_next
                    offs.append(pos - int(jptr[nvar - nv + kv]) - (i - 1) * jv_stride)
            assert offs[0] == offs[1], (scheme, r, a, offs)
            if offs[0] is None:
                if expr != 0:
                    E.dropped.append(("J", r, a))
                continue
            w = which(a)
            entry = (r, a, offs[0], expr)
            (E.j_own if w == "own" else E.j_next if w == "next" else E.j_v).append(entry)

    # Lagrangian of the step and its derivatives
    lagr = sum(E.lam[r] * rows[r] for r in range(CB)) + E.sg * E.cost
    grad = [D.d(lagr, a) for a in range(L.nz)]
    for b in range(L.nz):
        for a in range(b, L.nz):
            expr = D.d(grad[a], b)
            wa, wb = which(a), which(b)
            if wb == "v":
                kv1, kv2 = a - L.n_own - L.n_next, b - L.n_own - L.n_next
                in_pattern = pat.find_pos(hr, hptr, nvar - nv + kv1 + 1, nvar - nv + kv2 + 1) >= 0
                if expr != 0 and in_pattern:
                    E.h_vv.append((kv1, kv2, kv1 * (kv1 + 1) // 2 + kv2, expr))
                elif expr != 0:
                    E.dropped.append(("H", a, b))
                continue
            offs = []
            for i in (PROBE_I, PROBE_I + 1):
                pos = pat.find_pos(hr, hptr, gcol(i, a), gcol(i, b))
                if pos < 0:
                    offs.append(None)
                else:
                    offs.append(pos - (hbase(i) if wb == "own" else hbase(i + 1)))
            assert offs[0] == offs[1], (scheme, a, b, offs)
            if offs[0] is None:
                if expr != 0:
                    E.dropped.append(("H", a, b))
                continue
            if wb == "own":
                E.h_own.append((a, b, offs[0], expr))       # owner writes every slot, zeros included
            elif expr != 0:
                E.h_next.append((a, b, offs[0], expr))      # accumulated into the next block
    E.h_vv_table = [(k1, k2, k1 * (k1 + 1) // 2 + k2) for k1 in range(nv) for k2 in range(k1 + 1)
                    if pat.find_pos(hr, hptr, nvar - nv + k1 + 1, nvar - nv + k2 + 1) >= 0]
    # coverage: every slot of an interior column block must be written by its owner / the previous step
    jslots = sorted([e[2] for e in E.j_own] + [e[2] for e in E.j_next])
    assert jslots == list(range(E.j_block)), (scheme, jslots, E.j_block)
    hslots = sorted(e[2] for e in E.h_own)
    assert hslots == list(range(E.h_block)), (scheme, hslots, E.h_block)
    if nv > 0:
        for kv in range(nv):
            # (shifted rows sit in the next step's chunk of the column: slot >= jv_stride)
            sl = sorted(e[2] % E.jv_stride for e in E.j_v if e[1] - L.n_own - L.n_next == kv)
            assert sl == list(range(E.jv_stride)), (scheme, sl, E.jv_stride)

    # objective gradient of the running cost
    if ocp.lagrange is not None:
        for a in range(L.nz):
            expr = D.d(E.cost, a)
            w = which(a)
            if w == "own":
                E.g_own.append((a, a, a, expr))
            elif expr != 0:
                if w == "next":
                    E.g_next.append((a, a, a - L.n_own, expr))
                else:
                    E.g_v.append((a, a, a - L.n_own - L.n_next, expr))
    return E


def _fpath_symbols(ocp: OCP, scheme: str, L, ts, p):
    """Local variables and rows of the final-path element (see build_final_path)."""
    d = L.dims
    nx, nu, nv = d.nx, d.nu, d.nv
    cb = L.control_block
    nloc = cb + nx + nv
    z = [sp.Symbol(f"z{a}") for a in range(nloc)]
    v = z[cb + nx:]
    if scheme == "trapeze":
        x, ublk = z[0:nx], z[nx:nx + nu]
    else:
        ublk, x = z[0:cb], z[cb:cb + nx]
    if L.stagewise:
        _, Bw, _ = butcher_symbols(L.stage)
        u = [sum(Bw[j] * ublk[j * nu + k] for j in range(L.stage)) for k in range(nu)]
    else:
        u = ublk[:nu]      # u(tf) = first control of the last step (common.jl:140-155)
    rows = [sp.sympify(e) for e in ocp.path(ts.ti, x, u, v, p)] if d.np else []
    return z, rows


def _fpath_gcol(L, scheme, a):
    """1-based global column of local variable a of the final-path element."""
    d, N, B = L.dims, L.N, L.step_block
    cb = L.control_block
    nl = cb + d.nx
    if a >= nl:
        return L.nvar - d.nv + (a - nl) + 1
    if scheme == "trapeze":
        return N * B + a + 1
    if a < cb:
        return (N - 1) * B + d.nx + a + 1
    return N * B + (a - cb) + 1


def _bc_symbols(ocp: OCP, L, p):
    d = L.dims
    nx, nv = d.nx, d.nv
    nloc = 2 * nx + nv
    z = [sp.Symbol(f"z{a}") for a in range(nloc)]
    x0, xf, v = z[0:nx], z[nx:2 * nx], z[2 * nx:]
    rows = [sp.sympify(e) for e in ocp.boundary(x0, xf, v, p)] if d.nb else []
    cost = sp.sympify(ocp.mayer(x0, xf, v, p)) if ocp.mayer else sp.Integer(0)
    return z, rows, cost


def _bc_gcol(L, a):
    d, N, B = L.dims, L.N, L.step_block
    if a < d.nx:
        return a + 1
    if a < 2 * d.nx:
        return N * B + (a - d.nx) + 1
    return L.nvar - d.nv + (a - 2 * d.nx) + 1


def detect_structure(ocp: OCP, scheme: str, L, E):
    """Structural (":optimized") sparsity of the three element kinds, by tracing (gen/trace.py)."""
    d = L.dims
    nvo = L.n_own + L.n_next
    env = tr.make_env(E.z, E.ts, nvo, d.nv, E.inter)
    step_J, lag = [], tr.EMPTY
    for r, e in enumerate(E.rows_traced):
        t = tr.trace_expr(e, env)
        step_J += [(r, a) for a in sorted(t[0])]
        lag = tr._join(lag, t)
    lag = tr._join(lag, tr.trace_expr(E.cost_traced, env))
    step_H = sorted(tuple(sorted(q)) if len(q) == 2 else (min(q), min(q)) for q in lag[1])
    # final path element
    zf, rows_f = _fpath_symbols(ocp, scheme, L, E.ts, E.p)
    envf = tr.make_env(zf, TimeSyms(E.ts.ti, E.ts.tp, E.ts.h, E.ts.dti, {}, {}), L.control_block + d.nx, d.nv)
    fp_J, lagf = [], tr.EMPTY
    for r, e in enumerate(rows_f):
        t = tr.trace_expr(e, envf)
        fp_J += [(r, a) for a in sorted(t[0])]
        lagf = tr._join(lagf, t)
    fp_H = sorted(tuple(sorted(q)) if len(q) == 2 else (min(q), min(q)) for q in lagf[1])
    # boundary + Mayer element (no time dependence)
    zb, rows_b, cost_b = _bc_symbols(ocp, L, E.p)
    envb = {s: (frozenset([a]), frozenset()) for a, s in enumerate(zb)}
    bc_J, lagb = [], tr.EMPTY
    for r, e in enumerate(rows_b):
        t = tr.trace_expr(e, envb)
        bc_J += [(r, a) for a in sorted(t[0])]
        lagb = tr._join(lagb, t)
    lagb = tr._join(lagb, tr.trace_expr(cost_b, envb))
    bc_H = sorted(tuple(sorted(q)) if len(q) == 2 else (min(q), min(q)) for q in lagb[1])
    return dict(step_J=step_J, step_H=step_H, fp_J=fp_J, fp_H=fp_H, bc_J=bc_J, bc_H=bc_H)


def optimized_patterns_for(ocp: OCP, scheme: str, L, S):
    """CSC patterns of layout L from the detected structure S (detect_structure)."""
    d, N, cb = L.dims, L.N, L.cblock
    extra_J = [(N * cb + r + 1, _fpath_gcol(L, scheme, a)) for (r, a) in S["fp_J"]]
    extra_J += [(L.ncon - d.nb + r + 1, _bc_gcol(L, a)) for (r, a) in S["bc_J"]]
    extra_H = [(_fpath_gcol(L, scheme, a), _fpath_gcol(L, scheme, b)) for (a, b) in S["fp_H"]]
    extra_H += [(_bc_gcol(L, a), _bc_gcol(L, b)) for (a, b) in S["bc_H"]]
    shift_from = None
    if scheme == "euler_implicit" and d.np > 0:
        # shifted path rows; node 1 is the final-path structure on (U_1 block, x_1, v)
        shift_from = L.eqs_block
        cbk, nl = L.control_block, L.control_block + d.nx
        icol = lambda a: L.nvar - d.nv + (a - nl) + 1 if a >= nl else (d.nx + a + 1 if a < cbk else (a - cbk) + 1)
        extra_J += [(L.eqs_block + r + 1, icol(a)) for (r, a) in S["fp_J"]]
        extra_H += [(icol(a), icol(b)) for (a, b) in S["fp_H"]]
    return pat.optimized_patterns(L, S["step_J"], S["step_H"], extra_J, extra_H, shift_from)


@dataclass
class EdgeElement:
    """Final-path or boundary(+Mayer) element: dense local Jacobian / packed lower-triangular Hessian."""
    kind: str
    nloc: int
    nrows: int
    z: List[sp.Symbol]
    lam: List[sp.Symbol]
    rows: List[sp.Expr]
    cost: sp.Expr
    jac: list   # (r, a, r*nloc+a, expr) nonzero only
    hess: list  # (a, b, a*(a+1)/2+b, expr) nonzero only
    grad: list  # (a, a, a, expr) nonzero only (objective part)


def build_final_path(ocp: OCP, scheme: str, E: StepElement) -> EdgeElement:
    """g(t_{N+1}, x_{N+1}, u_f, v): DOCP_functions.jl:100,122-140.  Local order follows the global
    variable order: (u_f, x_f, v) when u(tf) aliases U_N (common.jl:144-146), (x_f, u_f, v) for trapeze."""
    L = E.L
    d = L.dims
    nx, nu, nv = d.nx, d.nu, d.nv
    cb = L.control_block
    nloc = cb + nx + nv
    z = [sp.Symbol(f"z{a}") for a in range(nloc)]
    lam = [sp.Symbol(f"lam{r}") for r in range(d.np)]
    v = z[cb + nx:]
    if scheme == "trapeze":
        x, ublk = z[0:nx], z[nx:nx + nu]
    else:
        ublk, x = z[0:cb], z[cb:cb + nx]
    if L.stagewise:
        _, Bw, _ = butcher_symbols(L.stage)
        u = [sum(Bw[j] * ublk[j * nu + k] for j in range(L.stage)) for k in range(nu)]
    else:
        u = ublk[:nu]
    ts = E.ts
    rows = [sp.sympify(e) for e in ocp.path(ts.ti, x, u, v, E.p)] if d.np else []
    D = Differ(z, cb + nx, nv, TimeSyms(ts.ti, ts.tp, ts.h, ts.dti, {}, {}))
    S = E.structure   # detected structure: its entries are listed even where the derivative vanishes
    jac, hess = [], []
    for r, e in enumerate(rows):
        for a in range(nloc):
            ex = D.d(e, a)
            if ex != 0 or (S and (r, a) in S["fp_J"]):
                jac.append((r, a, r * nloc + a, ex))
    lagr = sum(lam[r] * rows[r] for r in range(len(rows)))
    g = [D.d(lagr, a) for a in range(nloc)]
    for a in range(nloc):
        for b in range(a + 1):
            ex = D.d(g[a], b)
            if ex != 0 or (S and (b, a) in S["fp_H"]):
                hess.append((a, b, a * (a + 1) // 2 + b, ex))
    return EdgeElement("fpath", nloc, len(rows), z, lam, rows, sp.Integer(0), jac, hess, [])


def build_boundary(ocp: OCP, E: StepElement) -> EdgeElement:
    """b(x_1, x_{N+1}, v) rows (DOCP_functions.jl:103-111) + sigma * mayer(x_1, x_{N+1}, v) (l.35-39)."""
    d = E.L.dims
    nx, nv = d.nx, d.nv
    nloc = 2 * nx + nv
    z = [sp.Symbol(f"z{a}") for a in range(nloc)]
    lam = [sp.Symbol(f"lam{r}") for r in range(d.nb)]
    x0, xf, v = z[0:nx], z[nx:2 * nx], z[2 * nx:]
    rows = [sp.sympify(e) for e in ocp.boundary(x0, xf, v, E.p)] if d.nb else []
    cost = sp.sympify(ocp.mayer(x0, xf, v, E.p)) if ocp.mayer else sp.Integer(0)
    jac, hess, grad = [], [], []
    S = E.structure
    for r, e in enumerate(rows):
        for a in range(nloc):
            ex = sp.diff(e, z[a])
            if ex != 0 or (S and (r, a) in S["bc_J"]):
                jac.append((r, a, r * nloc + a, ex))
    lagr = sum(lam[r] * rows[r] for r in range(len(rows))) + E.sg * cost
    g = [sp.diff(lagr, z[a]) for a in range(nloc)]
    for a in range(nloc):
        for b in range(a + 1):
            ex = sp.diff(g[a], z[b])
            if ex != 0 or (S and (b, a) in S["bc_H"]):
                hess.append((a, b, a * (a + 1) // 2 + b, ex))
        ex = sp.diff(cost, z[a])
        if ex != 0:
            grad.append((a, a, a, ex))
    return EdgeElement("bc", nloc, len(rows), z, lam, rows, cost, jac, hess, grad)
