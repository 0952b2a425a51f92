"""Several controls per time step: the DirectShooting discretizer of the reference (src/direct_shooting.jl:55-71 builds
DOCP(ocp, grid_size, control_steps, :midpoint, nothing); the midpoint scheme handles control_steps > 1 at
src/ode/midpoint.jl:20,57-65,98-113,145-154; only exercised by test/manual_test.jl:20 in the reference).

Properties checked, as the reference has no stored numbers for this path:
  * layout: step block nx + control_steps * nu (midpoint.jl:20), bounds / initial guess on EVERY control of a step
    (DOCP_variables.jl:44,138), control array of N * control_steps + 1 rows and T_control (DOCP_data.jl:530,555-566);
  * with the controls of a step all equal, constraints and objective of control_steps = s coincide with the plain
    midpoint rule (the s sub-steps of length h/s all use the dynamics at the midpoint of the WHOLE step);
  * the detected pattern is the only one (direct_shooting.jl:101-137 hands no :manual pattern to ADNLPModels);
  * CUDA path == oracle (GPU)."""
import numpy as np
import pytest

from conftest import assert_close, column_scale, hess_scale, synthetic_point

CASES = [("goddard", 2), ("goddard", 3), ("goddard", 10), ("goddard_all", 2), ("double_integrator_minenergy", 10),
         ("double_integrator_freet0tf", 3), ("robot", 2), ("vanderpol", 10), ("bolza_freetf", 3)]


def _replicated(o, m, xm):
    """NLP vector of the control-steps layout whose controls of a step all equal the midpoint model's control."""
    cs, N = o.control_steps, o.N
    B, Bm = o.nx + cs * o.nu, m.nx + m.nu
    x = np.zeros(o.nvar)
    for i in range(N):
        x[i * B:i * B + o.nx] = xm[i * Bm:i * Bm + m.nx]
        for j in range(cs):
            x[i * B + o.nx + j * o.nu:i * B + o.nx + (j + 1) * o.nu] = xm[i * Bm + m.nx:i * Bm + m.nx + m.nu]
    x[N * B:] = xm[N * Bm:]
    return x


@pytest.mark.parametrize("problem,cs", CASES)
def test_oracle_layout_and_reduction_to_midpoint(oracle_mod, problem, cs):
    N = 9
    o = oracle_mod.OracleModel(problem, f"midpoint_cs{cs}", N, pattern_mode="optimized")
    m = oracle_mod.OracleModel(problem, "midpoint", N, pattern_mode="optimized")
    assert o.control_steps == cs and m.control_steps == 1
    assert o.step_block == o.nx + cs * o.nu                                   # midpoint.jl:20
    assert o.nvar == N * o.step_block + o.nx + o.nv and o.ncon == m.ncon      # midpoint.jl:25-30
    rng = np.random.default_rng(3)
    xm = m.x0() * (1 + 1e-2 * rng.uniform(-1, 1, m.nvar))
    x = _replicated(o, m, xm)
    assert_close(o.cons(x), m.cons(xm), scale=max(1.0, np.abs(xm).max()), rtol=1e-14, what="replicated controls cons")
    # the running cost samples the sub-step midpoints t_ij (midpoint.jl:106): equal only when it does not depend on t
    assert_close(o.obj(x), m.obj(xm), scale=1.0, rtol=1e-13, what="replicated controls obj")
    # bounds and the initial guess cover every control of a step
    lv, uv, lc, uc = o.bounds()
    lvm, uvm, lcm, ucm = m.bounds()
    assert np.array_equal(lv, _replicated(o, m, lvm)) and np.array_equal(uv, _replicated(o, m, uvm))
    assert np.array_equal(lc, lcm) and np.array_equal(uc, ucm)
    assert np.array_equal(o.x0(), _replicated(o, m, m.x0()))
    # every control of a step enters its state equation: the column of each control holds nx-row dependencies
    r, c = o.jac_structure()
    for j in range(cs):
        for k in range(o.nu):
            col = 2 * o.step_block + o.nx + j * o.nu + k + 1
            rows = r[c == col]
            assert len(rows) > 0 and rows.min() > 2 * o.cblock and rows.max() <= 3 * o.cblock, (j, k, rows)
    # unpacking: N * cs + 1 control rows, the last one the FIRST control of the last step (common.jl:97,140-155)
    y = rng.standard_normal(o.ncon)
    u = o.unpack(x, y)
    assert u["control"].shape == (N * cs + 1, o.nu) and u["control_time"].shape == (N * cs + 1,)
    B = o.step_block
    assert np.array_equal(u["control"][-1], x[(N - 1) * B + o.nx:(N - 1) * B + o.nx + o.nu])
    assert np.array_equal(u["control"][cs + 1], x[B + o.nx + o.nu:B + o.nx + 2 * o.nu])
    t = u["time"]
    assert u["control_time"][0] == t[0] and u["control_time"][-1] == t[-1]
    assert u["control_time"][cs] == t[1]
    assert np.all(np.diff(u["control_time"]) * np.sign(t[-1] - t[0]) > 0)      # free times: tf may lie below t0 at a trial point


@pytest.mark.parametrize("problem,cs", [("goddard", 2), ("double_integrator_minenergy", 10), ("robot", 2)])
def test_oracle_derivatives_by_central_differences(oracle_mod, problem, cs):
    o = oracle_mod.OracleModel(problem, f"midpoint_cs{cs}", 4, pattern_mode="optimized")
    rng = np.random.default_rng(5)
    x = o.x0() * (1 + 1e-2 * rng.uniform(-1, 1, o.nvar))
    y = rng.standard_normal(o.ncon)
    g, j, h = o.derivs(x, y, 0.7)
    r, c = o.jac_structure()
    J = np.zeros((o.ncon, o.nvar)); J[r - 1, c - 1] = j
    hr, hc = o.hess_structure()
    H = np.zeros((o.nvar, o.nvar)); H[hr - 1, hc - 1] = h
    H = H + np.tril(H, -1).T

    def lag_grad(xx):
        gg, jj, _ = o.derivs(xx, y, 0.7, want=("grad", "jac"))
        JJ = np.zeros((o.ncon, o.nvar)); JJ[r - 1, c - 1] = jj
        return 0.7 * gg + JJ.T @ y

    eps = 1e-6
    Jfd, Hfd = np.zeros_like(J), np.zeros_like(H)
    for k in range(o.nvar):
        e = np.zeros(o.nvar); e[k] = eps * max(1.0, abs(x[k]))
        Jfd[:, k] = (o.cons(x + e) - o.cons(x - e)) / (2 * e[k])
        Hfd[:, k] = (lag_grad(x + e) - lag_grad(x - e)) / (2 * e[k])
    # the detected pattern holds every nonzero of the dense derivative
    assert np.abs(J - Jfd).max() <= 1e-6 * max(1.0, np.abs(J).max())
    assert np.abs(H - Hfd).max() <= 1e-5 * max(1.0, np.abs(H).max())


def test_library_structure_and_host_queries(ctd, oracle_mod):
    """No GPU needed: dimensions, structure, bounds, initial guess through the C ABI against the oracle."""
    for problem, cs in CASES:
        N = 11
        o = oracle_mod.OracleModel(problem, f"midpoint_cs{cs}", N, pattern_mode="optimized")
        d = ctd.DOCP(problem, N, "midpoint", control_steps=cs)
        assert d.control_steps == cs and d.pattern_mode == "optimized"
        assert (d.nvar, d.ncon, d.nnzj, d.nnzh) == (o.nvar, o.ncon, o.nnzj, o.nnzh), (problem, cs)
        assert d.step_variables_block == o.nx + cs * o.nu
        for a, b in zip(d.jacobian_structure() + d.hessian_structure(), o.jac_structure() + o.hess_structure()):
            assert np.array_equal(a, b), (problem, cs)
        for a, b in zip(d.bounds(), o.bounds()):
            assert np.array_equal(a, b), (problem, cs)
        assert np.array_equal(d.initial_guess(), o.x0())
        # the discretizer spelling of the reference
        d2 = ctd.discretize(problem, ctd.DirectShooting(grid_size=N, control_steps=cs))
        assert (d2.nvar, d2.nnzj, d2.control_steps) == (d.nvar, d.nnzj, cs)
        t = np.linspace(0.0, 2.0, N + 1) ** 1.5
        tc = d.control_time(t)
        assert tc.shape == (N * cs + 1,) and tc[0] == t[0] and tc[-1] == t[-1] and tc[cs] == t[1]
        assert tc[1] == t[0] + 1 * (t[1] - t[0]) / cs                        # DOCP_data.jl:562
    assert ctd.DOCP("goddard", 5, "midpoint").control_steps == 1
    with pytest.raises(ctd.CtdError, match="midpoint scheme only"):
        ctd.DOCP("goddard", 5, "trapeze", control_steps=2)
    with pytest.raises(ctd.CtdError, match="not compiled in"):
        ctd.DOCP("goddard", 5, "midpoint", control_steps=7)


@pytest.mark.gpu
@pytest.mark.parametrize("N", [1, 2, 5, 130, 1000])
def test_cuda_path_matches_oracle(ctd, oracle_mod, N):
    sigma = 0.7
    for problem, cs in CASES:
        o = oracle_mod.OracleModel(problem, f"midpoint_cs{cs}", N, pattern_mode="optimized")
        d = ctd.DOCP(problem, N, "midpoint", control_steps=cs)
        nlp = d.nlp_model()
        x, y = synthetic_point(o, seed=77 + N)
        tag = f"{problem}/midpoint control_steps={cs}/N={N}"
        g_ref, j_ref, h_ref = o.derivs(x, y, sigma)
        sx = max(1.0, float(np.abs(x).max()))
        sj = column_scale(nlp.jac_structure(), j_ref, d.nvar)
        sh = hess_scale(o, x, y, sigma, h_ref)
        sg = 1e-3 * max(1.0, float(np.abs(g_ref).max()))
        assert_close(nlp.cons(x), o.cons(x), scale=sx, what=tag + " cons")
        assert_close(nlp.jac_coord(x), j_ref, scale=sj, what=tag + " jac")
        assert_close(nlp.hess_coord(x, y, obj_weight=sigma), h_ref, scale=sh, what=tag + " hess")
        r = nlp.eval_all(x, y, obj_weight=sigma)
        assert_close(r["cons"], o.cons(x), scale=sx, what=tag + " fused cons")
        assert_close(r["jac"], j_ref, scale=sj, what=tag + " fused jac")
        assert_close(r["hess"], h_ref, scale=sh, what=tag + " fused hess")
        assert_close(r["grad"], g_ref, scale=sg, what=tag + " fused grad")
        assert_close(r["obj"], o.obj(x), scale=1.0, what=tag + " fused obj")


@pytest.mark.gpu
def test_cuda_path_on_a_nonuniform_grid(ctd, oracle_mod):
    """The C ABI takes a time grid together with control_steps (the reference's DirectShooting passes none; the
    normalized grid of a DOCP built by hand may be anything increasing): h_i / control_steps sub-steps of every step."""
    rng = np.random.default_rng(6)
    grid = np.concatenate([[0.0], np.sort(rng.uniform(0.02, 0.98, 44)), [1.0]])
    for problem, cs in (("goddard", 3), ("double_integrator_minenergy", 10), ("goddard_all", 2)):
        o = oracle_mod.OracleModel(problem, f"midpoint_cs{cs}", 45, time_grid=grid, pattern_mode="optimized")
        d = ctd.DOCP(problem, 45, "midpoint", time_grid=grid, control_steps=cs)
        x, y = synthetic_point(o, seed=2)
        r = d.nlp_model().eval_all(x, y, obj_weight=0.8)
        g_ref, j_ref, h_ref = o.derivs(x, y, 0.8)
        tag = f"{problem}/midpoint control_steps={cs}/non-uniform grid"
        assert_close(r["cons"], o.cons(x), scale=max(1.0, float(np.abs(x).max())), what=tag + " cons")
        assert_close(r["jac"], j_ref, scale=column_scale(d.jacobian_structure(), j_ref, d.nvar), what=tag + " jac")
        assert_close(r["hess"], h_ref, scale=hess_scale(o, x, y, 0.8, h_ref), what=tag + " hess")
        assert_close(r["obj"], o.obj(x), scale=1.0, what=tag + " obj")


@pytest.mark.gpu
def test_cuda_unpack_pack_and_batches(ctd, oracle_mod):
    import torch
    rng = np.random.default_rng(9)
    for problem, cs in (("goddard_all", 2), ("goddard", 10), ("robot", 2)):
        N = 37
        o = oracle_mod.OracleModel(problem, f"midpoint_cs{cs}", N, pattern_mode="optimized")
        d = ctd.DOCP(problem, N, "midpoint", control_steps=cs)
        x = o.x0() * (1 + 1e-2 * rng.uniform(-1, 1, o.nvar))
        y = rng.standard_normal(o.ncon)
        ref, got = o.unpack(x, y), d.unpack(x, y)
        for k in got:
            assert got[k].shape == ref[k].shape and np.array_equal(got[k], ref[k]), (problem, cs, k)
        assert np.array_equal(d.control_time(got["time"]), ref["control_time"])
        st = rng.standard_normal((N + 1, o.nx))
        ct = rng.standard_normal((N + 1, o.nu))
        va = rng.standard_normal(o.nv) if o.nv else None
        assert np.array_equal(d.pack(st, ct, va), o.pack(st, ct, va)), (problem, cs)
    # a batch of instances on the device: every instance equals its own single-instance oracle
    B, N, cs = 5, 300, 3
    d = ctd.DOCP("goddard", N, "midpoint", control_steps=cs, batch=B)
    o = oracle_mod.OracleModel("goddard", f"midpoint_cs{cs}", N, pattern_mode="optimized")
    x, y = synthetic_point(d, seed=31, batch=B)
    r = d.evaluate(ctd.CONS | ctd.JAC | ctd.HESS | ctd.GRAD | ctd.OBJ, torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda(), 0.4)
    d.synchronize()
    for b in range(B):
        g_ref, j_ref, h_ref = o.derivs(x[b], y[b], 0.4)
        tag = f"batch instance {b} control_steps={cs}"
        assert_close(r["cons"][b].cpu().numpy(), o.cons(x[b]), scale=max(1.0, np.abs(x[b]).max()), what=tag + " cons")
        assert_close(r["jac"][b].cpu().numpy(), j_ref, scale=column_scale(d.jacobian_structure(), j_ref, d.nvar), what=tag + " jac")
        assert_close(r["hess"][b].cpu().numpy(), h_ref, scale=hess_scale(o, x[b], y[b], 0.4, h_ref), what=tag + " hess")
