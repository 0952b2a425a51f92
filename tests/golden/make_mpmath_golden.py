"""Golden Jacobian / Lagrangian-Hessian ENTRIES at 50 digits, independent of oracle, generator and kernels.

The discretised Goddard problem (reference test/problems/goddard.jl:7-49: F0 + u F1, Cd = 310, beta = 500, b = 2,
Tmax = 3.5; free tf, Mayer cost r(tf), 4 boundary rows) is written out here from the reference's formulas --
  midpoint      src/ode/midpoint.jl:47-72,124-156   c_i = x_{i+1} - (x_i + h_i f(t_mid, (x_i + x_{i+1})/2, u_i, v))
  Gauss-Legendre 3 stages, constant control   src/ode/irk.jl:116-122,236-308
                c_x = x_{i+1} - (x_i + h_i sum_j b_j k_j),  c_j = k_j - f(t_i + c_j h_i, x_i + h_i sum_l a_jl k_l, u_i, v)
  time grid     src/DOCP_data.jl:437-458            t_i = t0 + s_i (tf - t0), s = LinRange(0, 1, N + 1)
  layout        src/ode/common.jl:113-170           [X_1, U_1 (, K_1) ... X_N, U_N (, K_N), X_{N+1}, V]
  implicit Euler WITH path constraints (goddard_all, test/problems/goddard.jl:87-158: path rows v, u, r + v + m + u + tf)
                src/ode/euler.jl:59-72,79-104,140-159   c_i = x_{i+1} - (x_i + h_i f(t_{i+1}, x_{i+1}, U_i, v)); node i reads
                the control of block i-1 (u(]t_i, t_{i+1}]) = U_i, u(t_1) = U_1); rows [state_i, path_i]_i, path_{N+1}, boundary
  midpoint, 2 controls per step (DirectShooting)   src/ode/midpoint.jl:20,57-65,145-154
                x_next = x_i; x_next += (h_i / 2) f(t_mid, x_mid, U_ij, v) for j = 1, 2; c_i = x_{i+1} - x_next
-- as ONE sympy expression per constraint row for a grid of N = 3 steps; sympy differentiates the rows and the
Lagrangian sigma * f + y' c symbolically (dense, no sparsity assumption), and mpmath evaluates every entry with 50
significant digits at a float64 point taken as exact rationals (the Butcher tableau enters with the float64 values the
reference computes from its own expressions, so the golden function is the one the float64 codes evaluate).

    python tests/golden/make_mpmath_golden.py        # writes tests/golden/mpmath_goddard_n3.npz (a few minutes)

The file holds, per scheme: x, y, sigma, the dense Jacobian (ncon x nvar), the dense lower triangle of the Hessian
(nvar x nvar), the constraint values and the objective, each rounded once to float64 from the 50-digit result.
"""
import math
import os

import mpmath
import numpy as np
import sympy as sp

mpmath.mp.dps = 50
N = 3
NX, NU = 3, 1
CD, BETA, B, TMAX = 310, 500, 2, sp.Rational(7, 2)


def dynamics(x, u):
    r, v, m = x
    D = CD * v**2 * sp.exp(-BETA * (r - 1))
    return [v, -D / m - 1 / r**2 + u * TMAX / m, -B * TMAX * u]     # F0(x) + u F1(x), goddard.jl:7-15


def exact(f):
    return sp.Rational(float(f))     # the float64 value, exactly


def build(scheme):
    if scheme == "midpoint":
        blk, kstages = NX + NU, 0
    elif scheme == "goddard_all|euler_implicit":
        blk, kstages = NX + NU, 0
    elif scheme == "midpoint_cs2":
        blk, kstages = NX + 2 * NU, 0
    else:
        s15 = math.sqrt(15)
        A = [[5.0 / 36.0, 2 / 9 - s15 / 15, 5 / 36 - s15 / 30],
             [5.0 / 36.0 + s15 / 24, 2.0 / 9.0, 5.0 / 36.0 - s15 / 24],
             [5 / 36 + s15 / 30, 2 / 9 + s15 / 15, 5.0 / 36.0]]            # irk.jl:116-122, evaluated in float64
        Bw = [5.0 / 18.0, 4.0 / 9.0, 5.0 / 18.0]
        A = [[exact(a) for a in row] for row in A]
        Bw = [exact(b) for b in Bw]
        blk, kstages = NX + NU + 3 * NX, 3
    nvar = N * blk + NX + 1
    X = sp.symbols(f"X0:{nvar}")
    tf = X[-1]
    sgrid = [exact(j / N) for j in range(N + 1)]
    t = [0 + s * (tf - 0) for s in sgrid]
    xs = lambda i: list(X[i * blk:i * blk + NX])                 # state at node i (0-based)
    us = lambda i: X[i * blk + NX]
    rows = []
    for i in range(N):
        h = t[i + 1] - t[i]
        xi, xn, u = xs(i), xs(i + 1), us(i)
        if scheme == "midpoint":
            f = dynamics([(a + b) / 2 for a, b in zip(xi, xn)], u)
            rows += [xn[k] - (xi[k] + h * f[k]) for k in range(NX)]
        elif scheme == "midpoint_cs2":
            xm = [(a + b) / 2 for a, b in zip(xi, xn)]
            hs = h / 2
            x_next = list(xi)
            for j in range(2):
                f = dynamics(xm, X[i * blk + NX + j])
                x_next = [x_next[k] + hs * f[k] for k in range(NX)]
            rows += [xn[k] - x_next[k] for k in range(NX)]
        elif scheme == "goddard_all|euler_implicit":
            f = dynamics(xn, u)                                   # at node i+1, with the control of block i
            rows += [xn[k] - (xi[k] + h * f[k]) for k in range(NX)]
            uc = us(max(i - 1, 0))                                # control read by node i: block i-1, block 1 at the first node
            rows += [xi[1], uc, xi[0] + xi[1] + xi[2] + uc + tf]  # path rows of node i
        else:
            K = [list(X[i * blk + NX + NU + j * NX:i * blk + NX + NU + (j + 1) * NX]) for j in range(3)]
            rows += [xn[k] - (xi[k] + h * sum(Bw[j] * K[j][k] for j in range(3))) for k in range(NX)]
            for j in range(3):
                xij = [xi[k] + sum(h * A[j][l] * K[l][k] for l in range(3)) for k in range(NX)]
                f = dynamics(xij, u)
                rows += [K[j][k] - f[k] for k in range(NX)]
    x0, xf = xs(0), xs(N)
    if scheme == "goddard_all|euler_implicit":                    # path rows of node N+1: u(tf) = U_N
        uc = us(N - 1)
        rows += [xf[1], uc, xf[0] + xf[1] + xf[2] + uc + tf]
    rows += [x0[0], x0[1], x0[2], xf[2]]                         # boundary rows, goddard.jl:26-31
    obj = xf[0]                                                  # Mayer r(tf) (max problem: not negated, collocation.jl:145)
    return X, rows, obj


def point(nvar, ncon, blk, seed):
    rng = np.random.default_rng(seed)
    x = np.full(nvar, 0.1)
    for i in range(N + 1):
        x[i * blk:i * blk + NX] = [1.01, 0.05, 0.8]              # init_state of the problem, goddard.jl:48
    x = x * (1.0 + 1e-3 * rng.uniform(-1.0, 1.0, nvar))
    return x, rng.standard_normal(ncon)


ALL_SCHEMES = ("midpoint", "gauss_legendre_3_constant_control", "goddard_all|euler_implicit", "midpoint_cs2")


def main(only=None):
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "mpmath_goddard_n3.npz")
    out = dict(np.load(path)) if (only and os.path.exists(path)) else {}
    for scheme in (only or ALL_SCHEMES):
        X, rows, obj = build(scheme)
        nvar, ncon = len(X), len(rows)
        blk = {"gauss_legendre_3_constant_control": NX + NU + 3 * NX, "midpoint_cs2": NX + 2 * NU}.get(scheme, NX + NU)
        x, y = point(nvar, ncon, blk, 20261018)
        sigma = 0.7
        subs = {X[k]: exact(x[k]) for k in range(nvar)}
        ev = lambda e: float(sp.N(e.subs(subs), 50)) if e != 0 else 0.0
        Y = [exact(v) for v in y]
        lag = exact(sigma) * obj + sum(Y[r] * rows[r] for r in range(ncon))
        J = np.zeros((ncon, nvar))
        for r in range(ncon):
            for k in rows[r].free_symbols:
                J[r, X.index(k)] = ev(sp.diff(rows[r], k))
        H = np.zeros((nvar, nvar))
        grad = [sp.diff(lag, X[a]) for a in range(nvar)]
        for a in range(nvar):
            for k in grad[a].free_symbols:
                b = X.index(k)
                if b <= a:
                    H[a, b] = ev(sp.diff(grad[a], k))
        out[scheme + "|x"], out[scheme + "|y"], out[scheme + "|sigma"] = x, y, np.array(sigma)
        out[scheme + "|jac"], out[scheme + "|hess"] = J, H
        out[scheme + "|cons"] = np.array([ev(r) for r in rows])
        out[scheme + "|obj"] = np.array(ev(obj))
        print(scheme, "nvar", nvar, "ncon", ncon, "nnz J", int((J != 0).sum()), "nnz tril H", int((H != 0).sum()))
    np.savez_compressed(path, **out)


if __name__ == "__main__":
    import sys
    main(sys.argv[1:] or None)     # optional: only these schemes (the others are kept from the existing file)
