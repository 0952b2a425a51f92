"""Structural sparsity detection on sympy expressions, for the `:optimized` pattern mode.

The reference's default backend (`:optimized`, test/test_common.jl:33) lets ADNLPModels DETECT the Jacobian /
Hessian patterns by tracing the closures with SparseConnectivityTracer [3P].  Its global tracers classify
every operator by which of its first / second partial derivatives can be nonzero, without looking at values;
this module applies the same classification to the expression trees of the generator:
    a + b          dependencies of both, no new pair
    a * b          the cross pairs  dep(a) x dep(b)      (a constant factor is linear, even a structural zero)
    a ** n, f(a)   the pairs  dep(a) x dep(a)             (n != 1; f = exp, sin, cos, log, ...)
The reference pins the counts this produces for Goddard / midpoint / N = 250: nnzj 4504, nnzh 5259
(test/ci/test_modeler_solver.jl:32); the oracle re-derives them with an operator-overloading tracer
(oracle/tracer.hpp), independently of this code.
"""
from __future__ import annotations

import sympy as sp

EMPTY = (frozenset(), frozenset())


def _pairs(ga, gb):
    return frozenset(frozenset((a, b)) for a in ga for b in gb)


def _join(x, y):
    return (x[0] | y[0], x[1] | y[1])


def _nonlinear(x):
    return (x[0], x[1] | _pairs(x[0], x[0]))


def trace_expr(e, env):
    """(first-order set, set of unordered pairs) of expression `e`; `env` maps symbols to their traces
    (symbols that are absent are constants)."""
    e = sp.sympify(e)
    if e.is_Number or e.is_NumberSymbol:
        return EMPTY
    if e.is_Symbol:
        return env.get(e, EMPTY)
    if e.is_Add:
        r = EMPTY
        for a in e.args:
            r = _join(r, trace_expr(a, env))
        return r
    if e.is_Mul:
        r = None
        for a in e.args:
            t = trace_expr(a, env)
            if not t[0]:
                continue            # numeric / parameter factor: linear
            if r is None:
                r = t
            else:
                r = (r[0] | t[0], r[1] | t[1] | _pairs(r[0], t[0]))
        return r if r is not None else EMPTY
    if e.is_Pow:
        b, x = e.args
        tb, tx = trace_expr(b, env), trace_expr(x, env)
        if not tx[0]:
            if x == 1:
                return tb
            return _nonlinear(tb)
        j = _join(tb, tx)
        return (j[0], j[1] | _pairs(j[0], j[0]))
    if isinstance(e, sp.Function) or e.is_Function:
        r = EMPTY
        for a in e.args:
            r = _join(r, trace_expr(a, env))
        return (r[0], r[1] | _pairs(r[0], r[0]))
    raise NotImplementedError(f"cannot trace {type(e)}: {e}")


def make_env(z, ts, nv_off, nv, inter=()):
    """Traces of the local variables, the time symbols (linear in the free times) and the intermediates."""
    env = {s: (frozenset([a]), frozenset()) for a, s in enumerate(z)}
    free = sorted(set(ts.dti) | set(ts.dtp) | set(ts.dh))
    tdep = (frozenset(nv_off + k for k in free if k < nv), frozenset())
    for s in (ts.ti, ts.tp, ts.h):
        env[s] = tdep
    for sym, expr in inter:
        env[sym] = trace_expr(expr, env)
    return env
