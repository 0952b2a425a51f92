import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle_mod():
    """CPU oracle (test infrastructure): built on demand with g++."""
    from oracle import oracle as orc
    orc.build()
    return orc


@pytest.fixture(scope="session")
def ctd():
    """The product package; the CUDA library must exist (no fallback)."""
    import ctdirect_b200
    if not os.path.exists(ctdirect_b200.LIB_PATH):
        from ctdirect_b200 import build
        build.build()
    return ctdirect_b200


def synthetic_point(model, seed=20261018, batch=1):
    """Synthetic evaluation point of SURVEY.md §8d: the reference initial guess perturbed by
    1 + 1e-3*U(-1,1), multipliers ~ N(0,1).  `model` needs nvar, ncon and x0()/initial_guess()."""
    rng = np.random.default_rng(seed)
    x0 = model.x0() if hasattr(model, "x0") else model.initial_guess()
    x = x0[None, :] * (1.0 + 1e-3 * rng.uniform(-1.0, 1.0, size=(batch, x0.size)))
    y = rng.standard_normal((batch, model.ncon))
    if batch == 1:
        return x[0].copy(), y[0].copy()
    return np.ascontiguousarray(x), np.ascontiguousarray(y)


def column_scale(structure, ref, n, symmetric=False):
    """Per-entry tolerance floor of SURVEY.md §7 for a sparse value array in COO form: the largest magnitude
    in the entry's COLUMN (for the lower-triangular Hessian: in the row or column of the full symmetric
    matrix that bears the entry's column index).  `structure` = (rows, cols), 1-based."""
    rows, cols = structure
    ref = np.abs(np.asarray(ref, dtype=np.float64))
    cmax = np.zeros(n + 1)
    np.maximum.at(cmax, cols, ref)
    if symmetric:
        np.maximum.at(cmax, rows, ref)
        return np.maximum(cmax[cols], cmax[rows])
    return cmax[cols]


def hess_scale(o, x, y, sigma, h_ref):
    """Per-entry tolerance floor for Lagrangian-Hessian values: the larger of the column scale and the sum of the
    MAGNITUDES of the contributions that are added up to form the entry, one contribution per constraint-row
    type of a step (state equations, stage equations, path constraints: `o.cblock` of them) and step parity,
    plus the boundary rows and the cost.  An entry is a sum over rows with multipliers of either sign: at a 3-stage Gauss-Legendre
    step of Goddard three stage contributions of -2249.4, -522.2 and +2772.8 leave 1.147, and 1e-12 of THAT is
    below one ulp of the terms.  `o` is the oracle model (test infrastructure)."""
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    terms = np.zeros_like(np.asarray(h_ref, dtype=np.float64))
    nstep = o.N * o.cblock
    for r in range(o.cblock):
        for parity in (0, 1):   # neighbouring steps add into the same entries of the block they share
            yy = np.zeros_like(y)
            sl = slice(r + parity * o.cblock, nstep, 2 * o.cblock)
            yy[sl] = y[sl]
            terms += np.abs(o.derivs(x, yy, 0.0, want=("hess",))[2])
    yy = np.zeros_like(y)
    yy[nstep:] = y[nstep:]
    terms += np.abs(o.derivs(x, yy, sigma, want=("hess",))[2])
    return np.maximum(column_scale(o.hess_structure(), h_ref, o.nvar, symmetric=True), terms)


WORST = {}   # kind ("jac", "hess", ...) -> (largest pure relative error seen, where)


def pytest_terminal_summary(terminalreporter):
    if WORST:
        terminalreporter.write_line("worst pure relative error |a-b|/max(|a|,|b|) over entries above 1e-300, by quantity:")
        for kind in sorted(WORST):
            terminalreporter.write_line(f"  {kind:10s} {WORST[kind][0]:.3e}   ({WORST[kind][1]})")


def assert_close(a, b, scale=None, rtol=1e-12, what=""):
    """|a-b| <= rtol * max(|a|, |b|, scale): relative 1e-12 with a floor at the natural scale of the
    quantities that were combined (entries of the :manual pattern can be structurally present but
    numerically zero or the result of cancellation).  For Jacobian / Hessian values `scale` is the
    per-entry column scale (column_scale above)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    if a.size == b.size == 1:
        a, b = a.reshape(1), b.reshape(1)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    if a.size == 0:
        return
    s = np.maximum(np.abs(a), np.abs(b))
    err = np.abs(a - b)
    big = s > 1e-300
    if big.any() and what:
        rel = float(np.max(err[big] / s[big]))
        kind = what.split()[-1]
        if rel > WORST.get(kind, (-1.0, ""))[0]:
            WORST[kind] = (rel, what)
    if scale is not None:
        s = np.maximum(s, scale)
    bad = ~(err <= rtol * s)
    if bad.any():
        k = int(np.argmax(np.where(bad, err / np.maximum(s, 1e-300), 0)))
        raise AssertionError(f"{what}: {int(bad.sum())} of {a.size} entries differ; worst at {k}: "
                             f"{a.flat[k]!r} vs {b.flat[k]!r} (scale {s.flat[k] if np.ndim(s) else s})")
