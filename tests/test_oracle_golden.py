"""The CPU oracle against every number the reference's own tests / archives pin for this path
(SURVEY.md §8c).  Citations are into the reference tree (CTDirect.jl v1.0.12)."""
import numpy as np
import pytest


def test_goddard_midpoint_manual_counts(oracle_mod):
    # test/ci/test_modeler_solver.jl:29-37: nnzj == 6028 && nnzh == 6519 for :manual, midpoint, N = 250
    m = oracle_mod.OracleModel("goddard", "midpoint", 250)
    assert (m.nnzj, m.nnzh) == (6028, 6519)
    assert (m.nvar, m.ncon) == (250 * 4 + 3 + 1, 250 * 3 + 4)


def test_goddard_midpoint_optimized_counts(oracle_mod):
    # test/ci/test_modeler_solver.jl:32 documents the DETECTED pattern of the default `:optimized` backend
    # (SparseConnectivityTracer over the closures): nnzj 4504, nnzh 5259 for Goddard / midpoint / N = 250.
    # The oracle re-derives it with an operator-overloading tracer over the restated functions.
    m = oracle_mod.OracleModel("goddard", "midpoint", 250, pattern_mode="optimized")
    assert (m.nnzj, m.nnzh) == (4504, 5259)
    assert (m.nnzj, m.nnzh) == (18 * 250 + 4, 21 * 250 + 9)        # SURVEY.md §8c closed forms


@pytest.mark.parametrize("N,nnzh_opt", [(1000, 11011), (10000, 110011)])
def test_goddard_all_trapeze_optimized_hessian_archive_counts(oracle_mod, N, nnzh_opt):
    # test/archives/AD_backend.md:57-64: "Hess nnz 11011 (opt)" at N = 1000, 110011 at N = 10000.
    # (The archived optimized Jacobian counts 28011 / 280011 were produced with the older F0 + u*F1
    # formulation of goddard_all, whose `u*0` terms add 2 detected entries per step; the functional
    # version in test/problems/goddard.jl:87-158 gives 26011 / 260011.)
    m = oracle_mod.OracleModel("goddard_all", "trapeze", N, pattern_mode="optimized")
    assert m.nnzh == nnzh_opt
    assert m.nnzj == 26 * N + 11


def test_optimized_pattern_contains_every_nonzero_derivative(oracle_mod):
    rng = np.random.default_rng(4)
    for prob, scheme in (("goddard", "midpoint"), ("goddard_all", "trapeze"), ("robot", "gauss_legendre_2"),
                         ("vanderpol", "gauss_legendre_3")):
        m = oracle_mod.OracleModel(prob, scheme, 7, pattern_mode="optimized")
        man = oracle_mod.OracleModel(prob, scheme, 7, pattern_mode="corrected")
        x = m.x0() * (1 + 1e-2 * rng.uniform(-1, 1, m.nvar))
        y = rng.standard_normal(m.ncon)
        m.derivs(x, y, 0.7)
        assert m.dropped == (0, 0), (prob, scheme, m.dropped)      # no nonzero value falls outside the pattern
        assert m.nnzj <= man.nnzj and m.nnzh <= man.nnzh


@pytest.mark.parametrize("N,nvar,ncon,nnzh,nnzj_ref,nnzj_corr", [
    (1000, 4005, 6007, 30024, 39043, 42043),      # test/archives/AD_backend.md:57-64
    (10000, 40005, 60007, 300024, 390043, 420043),
])
def test_goddard_all_trapeze_archive_counts(oracle_mod, N, nvar, ncon, nnzh, nnzj_ref, nnzj_corr):
    m = oracle_mod.OracleModel("goddard_all", "trapeze", N)
    assert (m.nvar, m.ncon, m.nnzh) == (nvar, ncon, nnzh)
    # the archived Jacobian count predates trapeze.jl:203 (quirk Q1): current code omits dyn x v
    assert m.nnzj == nnzj_ref
    mc = oracle_mod.OracleModel("goddard_all", "trapeze", N, pattern_mode="corrected")
    assert mc.nnzj == nnzj_corr and mc.nnzh == nnzh


@pytest.mark.parametrize("scheme,stage,block,eqs,nvar,ncon", [
    ("gauss_legendre_2", 2, 5, 3, 16, 11),   # test/ci/test_discretization_stagewise.jl:176-198
    ("gauss_legendre_3", 3, 7, 4, 22, 14),
])
def test_stagewise_known_answer(oracle_mod, scheme, stage, block, eqs, nvar, ncon):
    grid = np.array([0.0, 0.2, 0.6, 1.0])
    m = oracle_mod.OracleModel("stagewise_scalar", scheme, time_grid=grid)
    assert (m.nvar, m.ncon, m.step_block, m.cblock) == (nvar, ncon, block, eqs)
    assert (m.nx, m.nu, m.nv, m.nb) == (1, 1, 0, 2)
    lvar, uvar, lcon, ucon = m.bounds()
    # stage controls of step 1 carry the control box 0 <= u <= 2 (l.67-71)
    for j in range(stage):
        assert lvar[1 + j] == 0.0 and uvar[1 + j] == 2.0
    # exact trajectory x = t^2, u = k = 2t at the Gauss nodes (l.16-41)
    c_nodes = {2: [0.5 - np.sqrt(3) / 6, 0.5 + np.sqrt(3) / 6],
               3: [0.5 - 0.1 * np.sqrt(15), 0.5, 0.5 + 0.1 * np.sqrt(15)]}[stage]
    xu = np.zeros(nvar)
    for i in range(4):
        xu[i * block] = grid[i] ** 2
    for i in range(3):
        h = grid[i + 1] - grid[i]
        for j in range(stage):
            tij = grid[i] + c_nodes[j] * h
            xu[i * block + 1 + j] = 2 * tij            # stage control
            xu[i * block + 1 + stage + j] = 2 * tij    # stage variable k
    c = m.cons(xu)
    np.testing.assert_allclose(c, lcon, atol=1e-12)    # l.93-98
    np.testing.assert_allclose(c, ucon, atol=1e-12)
    assert abs(m.obj(xu) - 4.0 / 3.0) <= 1e-12         # l.99


def test_default_initial_guess_is_point_one(oracle_mod):
    # test/ci/test_initial_guess.jl:32-54: default guess 0.1 everywhere (here: a problem without init)
    m = oracle_mod.OracleModel("double_integrator_minenergy", "trapeze", 100)
    assert (m.nvar, m.ncon, m.nnzj, m.nnzh) == (303, 204, 1216, 1510)   # BASELINE config 1
    assert np.all(m.x0() == 0.1)


def test_zero_control_layout_counts(oracle_mod):
    # test/ci/test_zero_control_allocations.jl:31,138: midpoint, N=10, NLP_u = 0
    m = oracle_mod.OracleModel("estimate_initial_condition", "midpoint", 10)
    assert (m.nu, m.nvar) == (0, 10 * 2 + 2 + 2)
    m = oracle_mod.OracleModel("estimate_rotation_rate", "midpoint", 10)
    assert (m.nu, m.nv, m.nvar) == (0, 1, 10 * 2 + 2 + 1)
    lvar, uvar, _, _ = m.bounds()
    assert len(lvar) == m.nvar and lvar[-1] == 0.0 and uvar[-1] == 10.0


def test_time_grid_validation(oracle_mod):
    # src/DOCP_data.jl:187: non-increasing grids are rejected
    with pytest.raises(ValueError, match="not strictly increasing"):
        oracle_mod.OracleModel("goddard", "midpoint", time_grid=[0.0, 0.5, 0.5, 1.0])
    # normalisation (l.191-196): a grid on [1, 3] is mapped to [0, 1]
    m = oracle_mod.OracleModel("double_integrator_minenergy", "midpoint", time_grid=[1.0, 1.5, 3.0])
    x = m.x0()
    np.testing.assert_allclose(m.time_grid(x), [0.0, 0.5, 2.0])   # t0 + s*(tf - t0) with T = 2


def test_unknown_scheme(oracle_mod):
    with pytest.raises(ValueError, match="Unknown discretization method"):
        oracle_mod.OracleModel("goddard", "runge_kutta_4", 10)


def test_double_integrator_analytic_solution_residual(oracle_mod):
    # test/ci/test_modeler_solver.jl:53-57 with T = 2 rescaled: x* = (3t^2/4 - t^3/4, 3t/2 - 3t^2/4), u* = 3/2 - 3t/2
    # the exact trajectory satisfies the continuous ODE, so the discrete defects are O(h^2) and the
    # trapezoidal objective converges to the analytic cost  int u*^2 = 3/2
    for N, tol in ((100, 1e-3), (1000, 1e-5)):
        m = oracle_mod.OracleModel("double_integrator_minenergy", "trapeze", N)
        t = np.linspace(0.0, 2.0, N + 1)
        x = np.zeros(m.nvar)
        x[0::3][:N + 1] = 0.75 * t ** 2 - 0.25 * t ** 3
        x[1::3][:N + 1] = 1.5 * t - 0.75 * t ** 2
        x[2::3][:N + 1] = 1.5 - 1.5 * t
        c = m.cons(x)
        assert np.abs(c[:-4]).max() < tol
        np.testing.assert_allclose(c[-4:], [0, 0, 1, 0], atol=1e-12)
        assert abs(m.obj(x) - 1.5) < 10 * tol


def test_oracle_derivatives_against_finite_differences(oracle_mod):
    """Independent check of the hyper-dual derivatives: central differences of cons / Lagrangian gradient."""
    for prob, scheme in (("goddard", "midpoint"), ("goddard_all", "trapeze"), ("robot", "gauss_legendre_2"),
                         ("vanderpol", "gauss_legendre_3_constant_control")):
        m = oracle_mod.OracleModel(prob, scheme, 4, pattern_mode="corrected")
        rng = np.random.default_rng(3)
        x = m.x0() * (1 + 1e-2 * rng.uniform(-1, 1, m.nvar))
        y = rng.standard_normal(m.ncon)
        g, j, h = m.derivs(x, y, 0.7)
        assert m.dropped == (0, 0)
        rows, cols = m.jac_structure()
        J = np.zeros((m.ncon, m.nvar))
        J[rows - 1, cols - 1] = j
        hr, hc = m.hess_structure()
        H = np.zeros((m.nvar, m.nvar))
        H[hr - 1, hc - 1] = h
        H = H + np.tril(H, -1).T
        eps = 1e-6
        Jfd = np.zeros_like(J)
        Gl = lambda xx: 0.7 * m.derivs(xx, None, 1.0, want=("grad",))[0] + _jt(m, xx, y)
        Hfd = np.zeros_like(H)
        for k in range(m.nvar):
            e = np.zeros(m.nvar)
            e[k] = eps * max(1.0, abs(x[k]))
            Jfd[:, k] = (m.cons(x + e) - m.cons(x - e)) / (2 * e[k])
            Hfd[:, k] = (Gl(x + e) - Gl(x - e)) / (2 * e[k])
        np.testing.assert_allclose(J, Jfd, rtol=1e-5, atol=1e-6 * max(1.0, np.abs(J).max()))
        np.testing.assert_allclose(H, Hfd, rtol=1e-4, atol=1e-5 * max(1.0, np.abs(H).max()))


def _jt(m, x, y):
    rows, cols = m.jac_structure()
    j = m.jac_coord(x)
    out = np.zeros(m.nvar)
    np.add.at(out, cols - 1, j * y[rows - 1])
    return out


def test_solution_unpacking_conventions(oracle_mod):
    # src/ode/common.jl:7-104: costate = multipliers of the state equations; final control duplicated when
    # u(tf) aliases U_N (common.jl:144-146); path multipliers divided by the step length (DOCP_data.jl:596-602)
    m = oracle_mod.OracleModel("goddard_all", "midpoint", 4)
    x = np.arange(m.nvar, dtype=float) + 1.0
    y = -(np.arange(m.ncon, dtype=float) + 1.0)
    s = m.unpack(x, y)
    B, CB = m.step_block, m.cblock
    assert np.array_equal(s["state"], np.array([x[i * B:i * B + 3] for i in range(5)]))
    assert np.array_equal(s["control"][:4, 0], x[3::B][:4]) and s["control"][4, 0] == s["control"][3, 0]
    assert s["variable"][0] == x[-1]
    tf = x[-1]
    np.testing.assert_allclose(s["time"], np.linspace(0, 1, 5) * tf, rtol=1e-15)
    assert np.array_equal(s["costate"], np.array([y[i * CB:i * CB + 3] for i in range(4)]))
    h = tf / 4
    np.testing.assert_allclose(s["path_dual"][:4], np.array([y[i * CB + 3:i * CB + 6] for i in range(4)]) / h, rtol=1e-14)
    np.testing.assert_allclose(s["path_dual"][4], y[4 * CB:4 * CB + 3] / h, rtol=1e-14)
    assert np.array_equal(s["boundary_dual"], y[-4:])
    # stagewise: the reported control is the b-weighted average of the stage controls (irk_stagewise.jl:197-205)
    m = oracle_mod.OracleModel("stagewise_scalar", "gauss_legendre_2", 2)
    x = np.arange(m.nvar, dtype=float)
    s = m.unpack(x)
    assert s["control"][0, 0] == 0.5 * x[1] + 0.5 * x[2] and s["control"][2, 0] == s["control"][1, 0]


def _golden(name):
    import os
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name)


def test_reference_known_answers_file(oracle_mod):
    """tests/golden/reference_known_answers.json: the reference's own pinned numbers as data, each with its citation."""
    import json
    ka = json.load(open(_golden("reference_known_answers.json")))
    for e in ka["pattern_counts"] + ka["stagewise_dims"]:
        m = oracle_mod.OracleModel(e["problem"], e["scheme"], e["N"], pattern_mode=e.get("pattern_mode", "reference"))
        for key in ("nvar", "ncon", "nnzj", "nnzh"):
            if key in e:
                assert getattr(m, key) == e[key], (e["cite"], key)
    m = oracle_mod.OracleModel("goddard", "gauss_legendre_2_constant_control", 4)
    x0 = m.x0()
    assert np.any(x0 == ka["default_initial_guess"]["value"])      # stage variables keep the default


def test_oracle_reproduces_regression_vectors(oracle_mod):
    """tests/golden/regression_vectors.npz (made by tests/golden/make_regression_vectors.py from this oracle: regression
    anchors, NOT reference outputs).  The oracle must still produce them bit for bit (same compiler flags, no
    reassociation) -- a drift here means the checker itself changed."""
    z = np.load(_golden("regression_vectors.npz"))
    tags = sorted({k.rsplit("|", 1)[0] for k in z.files})
    assert len(tags) >= 9
    for tag in tags:
        p, s, n, mode = tag.split("|")
        o = oracle_mod.OracleModel(p, s, int(n), pattern_mode=mode)
        x, y = z[tag + "|x"], z[tag + "|y"]
        g, j, h = o.derivs(x, y, 0.75)
        for key, val in (("cons", o.cons(x)), ("jac", j), ("hess", h), ("grad", g), ("obj", np.array([o.obj(x)]))):
            ref = z[tag + "|" + key]
            assert val.shape == ref.shape, (tag, key)
            assert np.allclose(val, ref, rtol=1e-14, atol=0.0) or np.array_equal(val, ref), (tag, key)
