"""Host-side logic of the product library (no GPU needed): dimensions, sparsity structures, bounds and
initial guess through the C ABI, bit-exact against the CPU oracle; error behaviour."""
import numpy as np
import pytest


def test_every_compiled_pair_matches_oracle_structure(ctd, oracle_mod):
    models = ctd.list_models()
    assert len(models) >= 70
    for (p, s, m) in models:
        for N in (1, 2, 3, 4, 9, 40):
            d = ctd.DOCP(p, N, s, pattern_mode=m)
            o = oracle_mod.OracleModel(p, s, N, pattern_mode=m)
            tag = f"{p}/{s}/{m}/N={N}"
            assert (d.nvar, d.ncon, d.nnzj, d.nnzh) == (o.nvar, o.ncon, o.nnzj, o.nnzh), tag
            for a, b in zip(d.jacobian_structure() + d.hessian_structure(), o.jac_structure() + o.hess_structure()):
                assert a.dtype == np.int64 and np.array_equal(a, b), tag
            for a, b in zip(d.bounds(), o.bounds()):
                assert np.array_equal(a, b), tag
            assert np.array_equal(d.initial_guess(), o.x0()), tag


@pytest.mark.parametrize("problem,scheme,N,nvar,ncon,nnzj,nnzh", [
    # BASELINE.json configs
    ("double_integrator_minenergy", "trapeze", 100, 303, 204, 1216, 1510),
    ("goddard", "midpoint", 100_000, 400_004, 300_004, 2_400_028, 2_600_019),
    ("goddard", "gauss_legendre_3_constant_control", 20_000, 260_004, 240_004, 3_480_028, 2_080_022),
    ("goddard", "gauss_legendre_3", 20_000, 300_004, 240_004, 3_840_028, 2_700_025),
    ("goddard", "midpoint", 1000, 4004, 3004, 24_028, 26_019),
    ("goddard", "midpoint", 250, 1004, 754, 6028, 6519),   # test/ci/test_modeler_solver.jl:37
])
def test_baseline_config_dimensions(ctd, problem, scheme, N, nvar, ncon, nnzj, nnzh):
    d = ctd.DOCP(problem, N, scheme)
    assert (d.nvar, d.ncon, d.nnzj, d.nnzh) == (nvar, ncon, nnzj, nnzh)


def test_detected_pattern_counts_through_the_library(ctd):
    # the reference documents the detected (`:optimized`) pattern of Goddard / midpoint / N = 250:
    # nnzj 4504, nnzh 5259 (test/ci/test_modeler_solver.jl:32); archive: Hess nnz 11011 / 110011 for
    # goddard_all / trapeze at N = 1000 / 10000 (test/archives/AD_backend.md:57-64)
    d = ctd.DOCP("goddard", 250, "midpoint", pattern_mode="optimized")
    assert (d.nnzj, d.nnzh) == (4504, 5259)
    assert ctd.DOCP("goddard_all", 1000, "trapeze", pattern_mode="optimized").nnzh == 11011
    assert ctd.DOCP("goddard_all", 10000, "trapeze", pattern_mode="optimized").nnzh == 110011
    # every problem has the detected pattern; a pair that is not compiled says so: with path constraints the implicit
    # Euler rule reads U_{i-1} at node i, which the reference's :manual pattern does not hold (euler.jl:66-69,214-216)
    assert ctd.DOCP("bolza_freetf", 10, "midpoint", pattern_mode="optimized").nnzj > 0
    assert ctd.DOCP("goddard_all", 10, "euler_implicit", pattern_mode="optimized").nnzj > 0
    with pytest.raises(ctd.CtdError, match="not compiled in"):
        ctd.DOCP("goddard_all", 10, "euler_implicit", pattern_mode="reference")


def test_large_grid_structure_is_csc_sorted(ctd, oracle_mod):
    d = ctd.DOCP("goddard", 20_000, "midpoint")
    o = oracle_mod.OracleModel("goddard", "midpoint", 20_000)
    for (r, c), (ro, co) in ((d.jacobian_structure(), o.jac_structure()), (d.hessian_structure(), o.hess_structure())):
        assert np.array_equal(r, ro) and np.array_equal(c, co)
        key = c.astype(np.int64) * (1 << 32) + r
        assert np.all(np.diff(key) > 0)          # column-major, rows ascending, no duplicates
    hr, hc = d.hessian_structure()
    assert np.all(hr >= hc)                       # lower triangle


def test_meta_mirrors_nlpmodels(ctd):
    docp = ctd.discretize("goddard", ctd.Collocation(grid_size=50, scheme=":midpoint"))
    nlp = docp.nlp_model()
    m = nlp.meta
    assert (m.nvar, m.ncon, m.nnzj, m.nnzh) == (204, 154, 24 * 50 + 28, 26 * 50 + 19)
    assert m.minimize is False                    # Goddard maximises r(tf): collocation.jl:145
    assert m.x0.shape == (204,) and m.lvar.shape == (204,) and m.lcon.shape == (154,)
    assert np.all(m.lcon[:150] == 0) and np.all(m.ucon[:150] == 0)
    np.testing.assert_array_equal(m.lcon[-4:], [1.0, 0.0, 1.0, 0.6])
    assert m.x0[-1] == 0.1 and m.x0[0] == 1.01
    docp2, nlp2 = ctd.direct_transcription("goddard", grid_size=50, disc_method="midpoint")   # legacy spelling
    assert nlp2.meta.nnzj == m.nnzj
    assert ctd.Collocation().grid_size == 250 and ctd.Collocation().scheme == "midpoint"     # collocation.jl:20-42


def test_explicit_time_grid(ctd, oracle_mod):
    grid = np.array([1.0, 1.25, 2.0, 2.5, 3.0])
    d = ctd.DOCP("vanderpol", scheme="trapeze", time_grid=grid)
    assert d.steps == 4
    np.testing.assert_allclose(d.normalized_time_grid(), (grid - 1.0) / 2.0)   # DOCP_data.jl:191-196
    d = ctd.DOCP("goddard", 7, "midpoint")
    assert np.array_equal(d.normalized_time_grid(), np.arange(8) / 7.0)        # collect(LinRange(0,1,N+1))


def test_errors(ctd):
    with pytest.raises(ctd.CtdError, match="Unknown discretization method"):   # DOCP_data.jl:343-348
        ctd.DOCP("goddard", 10, "runge_kutta_4")
    with pytest.raises(ctd.CtdError, match="unknown problem"):
        ctd.DOCP("no_such_ocp", 10, "midpoint")
    with pytest.raises(ctd.CtdError, match="not strictly increasing"):          # DOCP_data.jl:187
        ctd.DOCP("goddard", scheme="midpoint", time_grid=[0.0, 0.5, 0.4, 1.0])
    with pytest.raises(ctd.CtdError, match="grid_size"):
        ctd.DOCP("goddard", 0, "midpoint")
    with pytest.raises(ctd.CtdError, match="params"):
        ctd.DOCP("goddard", 10, "midpoint", params=[1.0, 2.0])


def test_no_cpu_fallback(ctd):
    """Without a CUDA device every evaluation entry point fails loudly (CUDA error code, message)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    d = ctd.DOCP("goddard", 10, "midpoint")
    nlp = d.nlp_model()
    x = d.initial_guess()
    for call in (lambda: nlp.obj(x), lambda: nlp.grad(x), lambda: nlp.cons(x), lambda: nlp.jac_coord(x),
                 lambda: nlp.hess_coord(x, np.zeros(d.ncon)), lambda: nlp.eval_all(x, np.zeros(d.ncon)),
                 lambda: d.norms(x)):
        with pytest.raises(ctd.CtdError, match="no CUDA device|CUDA"):
            call()


def test_stream_and_communicator_entry_points_without_a_device(ctd):
    """The newer entry points follow the same convention: host-only queries work anywhere (KKT structure, NCCL id),
    anything that needs the device returns a CUDA error instead of crashing."""
    import torch
    d = ctd.DOCP("goddard", 10, "midpoint")
    assert d.nnzk == d.nnzh + d.nnzj + d.ncon + (d.nvar - int(np.sum(np.equal(*d.hessian_structure()))))
    uid = ctd.DOCP.nccl_unique_id()
    assert isinstance(uid, bytes) and len(uid) == 128 and any(uid)
    with pytest.raises(ValueError):
        d.comm_init(b"short", 0, 1)
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    for call in (lambda: d.comm_init(uid, 0, 1), lambda: d.set_stream(None), lambda: d.comm_wait(0)):
        with pytest.raises(ctd.CtdError):
            call()


def test_product_never_touches_the_oracle():
    """No file of the product package imports, loads, links or executes anything under oracle/."""
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "ctdirect.jl_b200")
    bad = re.compile(r"(import\s+oracle|from\s+oracle|liboracle|oracle/|oracle\.py|orc_\w+\s*\(|#include\s+\"[^\"]*oracle)")
    for dirpath, _, files in os.walk(pkg):
        if "_build" in dirpath or os.path.basename(dirpath) in ("lib", "variants", "__pycache__"):
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h", ".jl")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                m = bad.search(txt)
                assert m is None or "oracle/tracer.hpp)" in txt[m.start() - 5:m.end() + 20], (dirpath, f, m.group(0))


class _Kkt:
    pass


def _kkt_reference(d, hvals=None, jvals=None, dx=None, dc=None):
    """Lower triangle of [[H + diag(dx), J'], [J, -diag(dc)]] in CSC order, structure by scipy from the COO
    structures, values accumulated entry by entry (explicit zeros stay in the structure)."""
    import scipy.sparse as sp
    n, m = d.nvar, d.ncon
    jr, jc = d.jacobian_structure()
    hr, hc = d.hessian_structure()
    rows = np.concatenate([hr - 1, jr - 1 + n, np.arange(n), n + np.arange(m)])
    cols = np.concatenate([hc - 1, jc - 1, np.arange(n), n + np.arange(m)])
    P = sp.coo_matrix((np.ones(len(rows)), (rows, cols)), shape=(n + m, n + m)).tocsc()
    P.sort_indices()
    K = _Kkt()
    K.indptr, K.indices, K.nnz = P.indptr.astype(np.int64), P.indices.astype(np.int64), P.nnz
    if hvals is not None:
        vals = np.concatenate([hvals, jvals, np.zeros(n) if dx is None else dx, np.zeros(m) if dc is None else -dc])
        kcols = np.repeat(np.arange(n + m, dtype=np.int64), np.diff(K.indptr))
        pos = np.searchsorted(kcols * (n + m) + K.indices, cols.astype(np.int64) * (n + m) + rows)
        K.data = np.zeros(K.nnz)
        np.add.at(K.data, pos, vals)     # diagonal: (0 + H_jj) + dx_j, the kernel's arithmetic
    return K


@pytest.mark.parametrize("problem,scheme,N,mode", [
    ("goddard", "midpoint", 40, "reference"), ("goddard", "midpoint", 40, "optimized"),
    ("goddard_all", "trapeze", 7, "reference"), ("goddard", "gauss_legendre_3", 5, "reference"),
    ("double_integrator_minenergy", "trapeze", 100, "reference"), ("robot", "trapeze", 6, "optimized"),
    ("goddard", "midpoint", 1, "reference"),
])
def test_kkt_structure_matches_scipy(ctd, problem, scheme, N, mode):
    """ctd_kkt_structure (host only) against scipy's CSC of the same lower-triangular block matrix; the detected
    patterns lack some Hessian diagonal entries, which the KKT structure must still provide."""
    d = ctd.DOCP(problem, N, scheme, pattern_mode=mode)
    colptr, rows = d.kkt_structure()
    K = _kkt_reference(d)
    assert colptr.dtype == np.int64 and rows.dtype == np.int64
    assert d.nnzk == K.nnz == len(rows)
    assert np.array_equal(colptr - 1, K.indptr)
    assert np.array_equal(rows - 1, K.indices)
    assert np.all(rows[colptr[:-1] - 1] == np.arange(1, d.nvar + d.ncon + 1))   # every column starts on its diagonal
