"""Groundwork for DESIGN.md 7, item 1 (node-centred trapeze element): a numerical proof, on the CPU, that the
trapeze DOCP of the reference (src/ode/trapeze.jl) can be assembled from NODE-local quantities only.

Today a thread owns a STEP and differentiates the dynamics at both of its nodes, so every node is differentiated
twice.  With h_i = ds_i * T(v) (T = tf - t0), the Lagrangian splits into node terms
    psi_j(w_j, v) = - T(v) * m_j . f(t_j(v), w_j, v) + nu_j . g(t_j(v), w_j, v) + sigma * T(v) * omega_j * l(t_j(v), w_j, v)
    m_j = (lambda_{j-1} ds_{j-1} + lambda_j ds_j) / 2,   omega_j = (ds_{j-1} + ds_j) / 2,   w_j = (x_j, u_j)
plus the linear terms lambda_i . (x_{i+1} - x_i) and the boundary / Mayer element, so that
  * the Hessian block (w_j, w_j), the cross block (w_j, v) and the node's share of (v, v) come from ONE Hessian of
    psi_j -- there is no (w_j, w_{j+1}) coupling and nothing to add between neighbours;
  * the whole Jacobian column block j (rows of c_{j-1}, c_j and the path rows of node j) comes from ONE Jacobian of f;
  * only the v-columns and the constraint values themselves need the neighbour's f (nx numbers per node).
This test builds exactly that with sympy from the generator's OCP descriptions and checks it against the oracle
(values within 1e-12, and zero outside the node-local blocks of the :manual pattern).  Nothing here is used by the
product yet."""
import os
import sys

import numpy as np
import pytest
import sympy as sp

from conftest import assert_close, synthetic_point

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _node_functions(ocp):
    nx, nu, nv, npth = ocp.nx, ocp.nu, ocp.nv, ocp.np
    x = sp.symbols(f"x0:{nx}")
    u = sp.symbols(f"u0:{max(nu, 1)}")[:nu]
    v = sp.symbols(f"v0:{max(nv, 1)}")[:nv]
    p = sp.symbols(f"p0:{max(ocp.npar, 1)}")[:ocp.npar]
    s, sig, om = sp.symbols("s sigma omega")
    m = sp.symbols(f"m0:{nx}")
    nu_ = sp.symbols(f"n0:{max(npth, 1)}")[:npth]
    t0, tf = ocp.time_expr(0, v, p), ocp.time_expr(1, v, p)
    T = tf - t0
    t = t0 + s * T
    f = [sp.sympify(e).subs(sp.Symbol("CTD_STRUCT_ZERO"), 0) for e in ocp.dynamics(t, x, u, v, p)]
    g = list(ocp.path(t, x, u, v, p)) if ocp.path else []
    lag = ocp.lagrange(t, x, u, v, p) if ocp.lagrange else sp.Integer(0)
    psi = -T * sum(mk * fk for mk, fk in zip(m, f)) + sum(nk * gk for nk, gk in zip(nu_, g)) + sig * T * om * lag
    wv = list(x) + list(u) + list(v)
    args = [x, u, v, p, s, sig, om, m, nu_]
    lam = lambda e: sp.lambdify(args, e, "numpy")
    return dict(f=lam(sp.Matrix(f)), fw=lam(sp.Matrix(f).jacobian(wv)), g=lam(sp.Matrix(g) if g else sp.Matrix([0])),
                gw=lam(sp.Matrix(g).jacobian(wv) if g else sp.zeros(1, len(wv))), T=lam(T), Tv=lam(sp.Matrix([T]).jacobian(list(v)) if nv else sp.zeros(1, 1)),
                l=lam(lag), lw=lam(sp.Matrix([lag]).jacobian(wv)), H=lam(sp.hessian(psi, wv)))


def _boundary_functions(ocp):
    nx, nv = ocp.nx, ocp.nv
    x0 = sp.symbols(f"a0:{nx}")
    xf = sp.symbols(f"b0:{nx}")
    v = sp.symbols(f"v0:{max(nv, 1)}")[:nv]
    p = sp.symbols(f"p0:{max(ocp.npar, 1)}")[:ocp.npar]
    beta = sp.symbols(f"be0:{max(ocp.nb, 1)}")[:ocp.nb]
    sig = sp.Symbol("sigma")
    b = list(ocp.boundary(x0, xf, v, p)) if ocp.boundary else []
    may = ocp.mayer(x0, xf, v, p) if ocp.mayer else sp.Integer(0)
    z = list(x0) + list(xf) + list(v)
    phi = sum(bk * e for bk, e in zip(beta, b)) + sig * may
    args = [x0, xf, v, p, beta, sig]
    lam = lambda e: sp.lambdify(args, e, "numpy")
    return dict(b=lam(sp.Matrix(b) if b else sp.Matrix([0])), bz=lam(sp.Matrix(b).jacobian(z) if b else sp.zeros(1, len(z))),
                may=lam(may), mayz=lam(sp.Matrix([may]).jacobian(z)), H=lam(sp.hessian(phi, z)))


@pytest.mark.parametrize("problem,N,grid", [
    ("goddard_all", 7, None),                                                   # free tf, path constraints, Mayer
    ("double_integrator_minenergy", 9, np.array([0, .05, .2, .25, .5, .6, .7, .85, .9, 1.0])),   # Lagrange, parameter tf
    ("robot", 4, None),                                                         # 6 states, 3 controls, min tf
])
def test_trapeze_from_node_local_quantities(oracle_mod, problem, N, grid):
    from ctdirect_b200.gen import ocps
    ocp = ocps.ALL[problem]
    o = oracle_mod.OracleModel(problem, "trapeze", N, time_grid=grid, pattern_mode="corrected")
    nx, nu, nv, npth, nb = ocp.nx, ocp.nu, ocp.nv, ocp.np, ocp.nb
    B, CB = nx + nu, nx + npth
    x, y = synthetic_point(o, seed=31)
    sigma = 0.7
    par = [val for _, val in ocp.params]
    sg = np.arange(N + 1) / N if grid is None else grid
    ds = np.diff(sg)
    F, Fb = _node_functions(ocp), _boundary_functions(ocp)
    vv = x[o.nvar - nv:]
    X = lambda j: x[(j - 1) * B:(j - 1) * B + nx]            # node j = 1..N+1 (trapeze stores a final control)
    U = lambda j: x[(j - 1) * B + nx:(j - 1) * B + B]
    lam = lambda i: y[(i - 1) * CB:(i - 1) * CB + nx] if 1 <= i <= N else np.zeros(nx)
    nuj = lambda j: y[(j - 1) * CB + nx:(j - 1) * CB + CB] if j <= N else y[N * CB:N * CB + npth]
    dsj = lambda i: ds[i - 1] if 1 <= i <= N else 0.0
    wcols = lambda j: list(range((j - 1) * B, j * B)) + list(range(o.nvar - nv, o.nvar))     # 0-based columns of (w_j, v)

    cons = np.zeros(o.ncon)
    J, H, grad = {}, {}, np.zeros(o.nvar)
    obj = 0.0
    add = lambda d, r, c, val: d.__setitem__((r, c), d.get((r, c), 0.0) + val)
    fs, Tval = {}, None
    for j in range(1, N + 2):
        m = 0.5 * (lam(j - 1) * dsj(j - 1) + lam(j) * dsj(j))
        om = 0.5 * (dsj(j - 1) + dsj(j))
        a = (X(j), U(j), vv, par, sg[j - 1], sigma, om, m, nuj(j) if npth else [])
        f = np.asarray(F["f"](*a), dtype=float).reshape(-1)
        fw = np.asarray(F["fw"](*a), dtype=float).reshape(nx, B + nv)
        Tval = float(F["T"](*a))
        Tv = np.asarray(F["Tv"](*a), dtype=float).reshape(-1)[:nv]
        fs[j] = f
        cols = wcols(j)
        # ---- Jacobian column block j and this node's share of the v columns
        for i, sgn in ((j - 1, +1.0), (j, -1.0)):            # rows of c_{j-1} (x_j enters with +) and c_j (with -)
            if not (1 <= i <= N):
                continue
            h = ds[i - 1] * Tval
            for k in range(nx):
                r = (i - 1) * CB + k
                add(J, r, (j - 1) * B + k, sgn)
                for q, c in enumerate(cols):
                    add(J, r, c, -0.5 * h * fw[k, q])
                for q in range(nv):                          # d h_i / d v * f_j / 2
                    add(J, r, o.nvar - nv + q, -0.5 * ds[i - 1] * Tv[q] * f[k])
        if npth:
            gw = np.asarray(F["gw"](*a), dtype=float).reshape(npth, B + nv)
            gval = np.asarray(F["g"](*a), dtype=float).reshape(-1)
            r0 = (j - 1) * CB + nx if j <= N else N * CB
            cons[r0:r0 + npth] = gval
            for k in range(npth):
                for q, c in enumerate(cols):
                    add(J, r0 + k, c, gw[k, q])
        # ---- Hessian: ONE Hessian of psi_j gives (w_j, w_j), (w_j, v) and the node's share of (v, v)
        Hn = np.asarray(F["H"](*a), dtype=float).reshape(B + nv, B + nv)
        for q1, c1 in enumerate(cols):
            for q2, c2 in enumerate(cols):
                if c1 >= c2:
                    add(H, c1, c2, Hn[q1, q2])
        # ---- running cost and its gradient (trapeze quadrature is node-local too: weight omega_j * T)
        if ocp.lagrange:
            lval = float(F["l"](*a))
            lw = np.asarray(F["lw"](*a), dtype=float).reshape(-1)
            obj += Tval * om * lval
            for q, c in enumerate(cols):
                grad[c] += Tval * om * lw[q]
            for q in range(nv):
                grad[o.nvar - nv + q] += Tv[q] * om * lval
    # ---- constraint values: the only place that needs the neighbour's f (reference association, trapeze.jl:131-136)
    for i in range(1, N + 1):
        cons[(i - 1) * CB:(i - 1) * CB + nx] = X(i + 1) - (X(i) + 0.5 * (ds[i - 1] * Tval) * (fs[i] + fs[i + 1]))
    # ---- boundary constraints + Mayer cost (unchanged element)
    beta = y[o.ncon - nb:]
    ab = (X(1), X(N + 1), vv, par, beta, sigma)
    zc = list(range(nx)) + list(range(N * B, N * B + nx)) + list(range(o.nvar - nv, o.nvar))
    if nb:
        cons[o.ncon - nb:] = np.asarray(Fb["b"](*ab), dtype=float).reshape(-1)
        bz = np.asarray(Fb["bz"](*ab), dtype=float).reshape(nb, len(zc))
        for k in range(nb):
            for q, c in enumerate(zc):
                add(J, o.ncon - nb + k, c, bz[k, q])
    Hb = np.asarray(Fb["H"](*ab), dtype=float).reshape(len(zc), len(zc))
    for q1, c1 in enumerate(zc):
        for q2, c2 in enumerate(zc):
            if c1 >= c2:
                add(H, c1, c2, Hb[q1, q2])
    obj += float(Fb["may"](*ab))
    mz = np.asarray(Fb["mayz"](*ab), dtype=float).reshape(-1)
    for q, c in enumerate(zc):
        grad[c] += mz[q]

    # ---- compare with the oracle on its (corrected :manual) pattern
    g_ref, j_ref, h_ref = o.derivs(x, y, sigma)
    jr, jc = o.jac_structure()
    hr, hc = o.hess_structure()
    jv = np.array([J.get((r - 1, c - 1), 0.0) for r, c in zip(jr, jc)])
    hv = np.array([H.get((r - 1, c - 1), 0.0) for r, c in zip(hr, hc)])
    tag = f"{problem}/trapeze/N={N}"
    assert_close(cons, o.cons(x), scale=max(1.0, np.abs(x).max()), what=tag + " cons")
    assert_close(jv, j_ref, scale=1e-3 * max(1.0, np.abs(j_ref).max()), what=tag + " jac")
    assert_close(hv, h_ref, scale=1e-3 * max(1e-300, np.abs(h_ref).max()), what=tag + " hess")
    assert_close(grad, g_ref, scale=1e-3 * max(1.0, np.abs(g_ref).max()), what=tag + " grad")
    assert_close(obj, o.obj(x), scale=1.0, what=tag + " obj")
    # every value the node-local assembly produced lies inside the oracle's pattern (nothing was dropped) ...
    inside_j = {(r - 1, c - 1) for r, c in zip(jr, jc)}
    inside_h = {(r - 1, c - 1) for r, c in zip(hr, hc)}
    assert all(abs(val) < 1e-300 or k in inside_j for k, val in J.items())
    assert all(abs(val) < 1e-300 or k in inside_h for k, val in H.items())
    # ... and the (w_j, w_{j+1}) blocks of the :manual pattern are structurally zero: no neighbour coupling
    cross = [k for k, (r, c) in enumerate(zip(hr, hc)) if r <= N * B + B and c <= N * B + B and (r - 1) // B != (c - 1) // B
             and not ((c - 1) // B == 0 and (r - 1) // B == N)]
    assert len(cross) > 0 and np.all(h_ref[cross] == 0.0)
