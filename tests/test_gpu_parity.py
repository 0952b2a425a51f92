"""Parity of the CUDA path (through the C ABI) with the CPU oracle.  Run on a B200: pytest -m gpu.

Tolerance (BASELINE.json north_star): indices bit-exact, values within 1e-12 relative.  `assert_close`
uses |a-b| <= 1e-12 * max(|a|, |b|, scale) where `scale` is, for Jacobian values, the largest magnitude in
the entry's COLUMN (entries of the :manual pattern can be structural zeros), for Hessian
values the larger of that column scale and the summed magnitudes of the per-row contributions that make up
the entry (conftest.hess_scale), for constraints the magnitude of x.  The largest PURE relative error met
by every comparison is printed at the end of the run."""
import os

import numpy as np
import pytest

from conftest import assert_close, column_scale, hess_scale, synthetic_point

pytestmark = pytest.mark.gpu

RTOL = 1e-12


def _models(ctd):
    return ctd.list_models()


def _scales(o, x, y, sigma):
    """Natural magnitudes: constraints combine entries of x; Jacobian entries are O(1) + h*df;
    Hessian entries are sums of |y|- and sigma-weighted second derivatives."""
    jref = o.jac_coord(x)
    sx = max(1.0, float(np.abs(x).max()))
    sj = max(1.0, float(np.abs(jref).max()))
    return sx, sj


@pytest.mark.parametrize("N", [1, 2, 3, 5, 130, 257])
def test_all_models_small_grids(ctd, oracle_mod, N):
    sigma = 0.7
    for (p, s, m) in _models(ctd):
        o = oracle_mod.OracleModel(p, s, N, pattern_mode=m)
        d = ctd.DOCP(p, N, s, pattern_mode=m)
        nlp = d.nlp_model()
        x, y = synthetic_point(o, seed=1234 + N)
        tag = f"{p}/{s}/{m}/N={N}"
        # structure: bit-exact
        for a, b in zip(nlp.jac_structure() + nlp.hess_structure(), o.jac_structure() + o.hess_structure()):
            assert np.array_equal(a, b), tag
        g_ref, j_ref, h_ref = o.derivs(x, y, sigma)
        c_ref, f_ref = o.cons(x), o.obj(x)
        sx = max(1.0, float(np.abs(x).max()))
        sj = column_scale(nlp.jac_structure(), j_ref, d.nvar)
        sh = hess_scale(o, x, y, sigma, h_ref)
        sg = max(1.0, float(np.abs(g_ref).max()))
        # separate callbacks
        assert_close(nlp.cons(x), c_ref, scale=sx, rtol=RTOL, what=tag + " cons")
        assert_close(nlp.jac_coord(x), j_ref, scale=sj, rtol=RTOL, what=tag + " jac")
        assert_close(nlp.hess_coord(x, y, obj_weight=sigma), h_ref, scale=sh, rtol=RTOL, what=tag + " hess")
        assert_close(nlp.obj(x), f_ref, scale=1.0, rtol=RTOL, what=tag + " obj")
        assert_close(nlp.grad(x), g_ref, scale=sg * 1e-3, rtol=RTOL, what=tag + " grad")
        # fused launch gives the same numbers as the separate ones
        r = nlp.eval_all(x, y, obj_weight=sigma)
        assert_close(r["cons"], c_ref, scale=sx, rtol=RTOL, what=tag + " fused cons")
        assert_close(r["jac"], j_ref, scale=sj, rtol=RTOL, what=tag + " fused jac")
        assert_close(r["hess"], h_ref, scale=sh, rtol=RTOL, what=tag + " fused hess")
        assert_close(r["grad"], g_ref, scale=sg * 1e-3, rtol=RTOL, what=tag + " fused grad")
        assert_close(r["obj"], f_ref, scale=1.0, rtol=RTOL, what=tag + " fused obj")


def _compare_all(ctd, o, d, x, y, sigma, tag, rtol=RTOL):
    nlp = d.nlp_model()
    g_ref, j_ref, h_ref = o.derivs(x, y, sigma)
    sh = hess_scale(o, x, y, sigma, h_ref)
    c_ref, f_ref = o.cons(x), o.obj(x)
    sx = max(1.0, float(np.abs(x).max()))
    r = nlp.eval_all(x, y, obj_weight=sigma)
    assert_close(r["cons"], c_ref, scale=sx, rtol=rtol, what=tag + " cons")
    assert_close(r["jac"], j_ref, scale=column_scale(nlp.jac_structure(), j_ref, d.nvar), rtol=rtol, what=tag + " jac")
    assert_close(r["hess"], h_ref, scale=sh, rtol=rtol, what=tag + " hess")
    assert_close(r["grad"], g_ref, scale=1e-3 * max(1.0, float(np.abs(g_ref).max())), rtol=rtol, what=tag + " grad")
    assert_close(r["obj"], f_ref, scale=1.0, rtol=rtol, what=tag + " obj")
    return r


@pytest.mark.parametrize("problem,scheme,N", [
    ("double_integrator_minenergy", "trapeze", 100),            # BASELINE config 1
    ("goddard", "midpoint", 100_000),                           # config 2
    ("goddard", "gauss_legendre_3_constant_control", 20_000),   # config 3 (constant control)
    ("goddard", "gauss_legendre_3", 20_000),                    # config 3 (stagewise = :gauss_legendre_3 today)
    ("robot", "trapeze", 100_000),                              # config 4 (synthetic robot arm)
])
def test_baseline_configs_full_size(ctd, oracle_mod, problem, scheme, N):
    for mode in ("reference", "corrected"):
        o = oracle_mod.OracleModel(problem, scheme, N, pattern_mode=mode)
        d = ctd.DOCP(problem, N, scheme, pattern_mode=mode)
        x, y = synthetic_point(o)
        r = _compare_all(ctd, o, d, x, y, 1.0, f"{problem}/{scheme}/{mode}/N={N}")
        # size-independent properties: the Hessian is linear in (y, obj_weight); evaluation is deterministic
        nlp = d.nlp_model()
        h1 = nlp.hess_coord(x, y, obj_weight=0.25)
        h2 = nlp.hess_coord(x, 2.0 * y, obj_weight=0.5)
        assert_close(h2, 2.0 * h1, scale=column_scale(nlp.hess_structure(), h1, d.nvar, symmetric=True), rtol=1e-13,
                     what="hessian linearity")
        r2 = nlp.eval_all(x, y, obj_weight=1.0)
        for k in ("cons", "jac", "hess", "grad", "obj"):
            assert np.array_equal(r[k], r2[k]), f"{k} not reproducible bit for bit"


@pytest.mark.parametrize("scheme", ["trapeze", "midpoint", "euler", "euler_implicit", "gauss_legendre_2", "gauss_legendre_3"])
def test_reference_scheme_sweep_problems(ctd, oracle_mod, scheme):
    """The two problems the reference sweeps over its six schemes (test/ci/test_discretization.jl:62-80): `beam`
    (Lagrange cost, state box) and `double_integrator_freet0tf` at grid_size 50 -- the only problem of the tree with
    a free INITIAL time, i.e. the `t = t0 + s (tf - t0)` seeds (1 - s) of both ends of a step."""
    for problem, N in (("beam", 250), ("double_integrator_freet0tf", 50)):
        for mode in ("reference", "corrected"):
            o = oracle_mod.OracleModel(problem, scheme, N, pattern_mode=mode)
            d = ctd.DOCP(problem, N, scheme, pattern_mode=mode)
            x, y = synthetic_point(o, seed=99)
            if problem == "double_integrator_freet0tf":
                x[-2:] = (0.3, 2.1)            # t0, tf well apart (the default guess 0.1, 0.1 has tf - t0 ~ 0)
            _compare_all(ctd, o, d, x, y, 0.7, f"{problem}/{scheme}/{mode}/N={N}")
            lv, uv, lc, uc = d.bounds()
            for a, b in zip((lv, uv, lc, uc), o.bounds()):
                assert np.array_equal(a, b)
            assert np.array_equal(d.initial_guess(), o.x0())
    assert ctd.DOCP("double_integrator_freet0tf", 50, scheme).free_t0


def test_detected_pattern_full_size(ctd, oracle_mod):
    """`:optimized` pattern at BASELINE size: same values as the manual pattern where both have an entry,
    every entry the manual pattern adds on top is a structural zero."""
    N = 100_000
    o = oracle_mod.OracleModel("goddard", "midpoint", N, pattern_mode="optimized")
    d = ctd.DOCP("goddard", N, "midpoint", pattern_mode="optimized")
    assert (d.nnzj, d.nnzh) == (18 * N + 4, 21 * N + 9)
    x, y = synthetic_point(o)
    r = _compare_all(ctd, o, d, x, y, 1.0, "goddard/midpoint/optimized/N=1e5")
    dm = ctd.DOCP("goddard", N, "midpoint")
    rm = dm.nlp_model().eval_all(x, y, obj_weight=1.0)
    for key, (ro, co), (rmn, cmn) in (("jac", d.jacobian_structure(), dm.jacobian_structure()),
                                      ("hess", d.hessian_structure(), dm.hessian_structure())):
        ko = co.astype(np.int64) * (1 << 32) + ro
        km = cmn.astype(np.int64) * (1 << 32) + rmn
        pos = np.searchsorted(km, ko)
        assert np.array_equal(km[pos], ko)                      # the detected pattern is a subset
        assert np.array_equal(rm[key][pos], r[key])             # same numbers, bit for bit
        rest = np.ones(len(km), dtype=bool)
        rest[pos] = False
        assert np.all(rm[key][rest] == 0.0)


def test_jacobian_directional_derivative_full_size(ctd):
    """J(x) d against central differences of c at grid size 1e5 (independent of the oracle)."""
    d = ctd.DOCP("goddard", 100_000, "midpoint")
    nlp = d.nlp_model()
    rng = np.random.default_rng(5)
    x = nlp.meta.x0 * (1 + 1e-3 * rng.uniform(-1, 1, d.nvar))
    dirv = rng.standard_normal(d.nvar)
    rows, cols = nlp.jac_structure()
    jd = np.zeros(d.ncon)
    np.add.at(jd, rows - 1, nlp.jac_coord(x) * dirv[cols - 1])
    eps = 1e-7
    fd = (nlp.cons(x + eps * dirv) - nlp.cons(x - eps * dirv)) / (2 * eps)
    assert np.abs(jd - fd).max() <= 1e-6 * max(1.0, np.abs(jd).max())


def test_batched_instances_match_single_instance_oracles(ctd, oracle_mod):
    """BASELINE config 5 in miniature and at full size: per-instance parameters (Tmax) and points."""
    rng = np.random.default_rng(11)
    for B, N, check in ((5, 64, range(5)), (4096, 1000, range(0, 4096, 517))):
        par = np.tile(np.array([310.0, 500.0, 2.0, 3.5]), (B, 1))
        par[:, 3] = rng.uniform(1.0, 3.5, B)      # continuation range, test/ci/test_initial_guess.jl:242
        d = ctd.DOCP("goddard", N, "midpoint", params=par, batch=B)
        x, y = synthetic_point(d, seed=77, batch=B)
        r = d.evaluate(ctd.CONS | ctd.JAC | ctd.HESS | ctd.GRAD | ctd.OBJ, x, y, 1.0)
        assert r["jac"].shape == (B, d.nnzj) and r["obj"].shape == (B,)
        obj, viol = d.norms(x)
        lvar, uvar, lcon, ucon = d.bounds()
        for b in check:
            o = oracle_mod.OracleModel("goddard", "midpoint", N, params=par[b])
            g, j, h = o.derivs(x[b], y[b], 1.0)
            c = o.cons(x[b])
            tag = f"instance {b}/{B}"
            assert_close(r["cons"][b], c, scale=max(1.0, np.abs(x[b]).max()), what=tag + " cons")
            assert_close(r["jac"][b], j, scale=column_scale(d.jacobian_structure(), j, d.nvar), what=tag + " jac")
            assert_close(r["hess"][b], h, scale=hess_scale(o, x[b], y[b], 1.0, h), what=tag + " hess")
            assert_close(r["grad"][b], g, scale=1e-3, what=tag + " grad")
            assert_close(r["obj"][b], o.obj(x[b]), scale=1.0, what=tag + " obj")
            assert_close(obj[b], o.obj(x[b]), scale=1.0, what=tag + " norms obj")
            v_ref = max(0.0, float(np.max(np.maximum(lcon - c, c - ucon))))
            assert_close(viol[b], v_ref, scale=max(1.0, np.abs(x[b]).max()), what=tag + " violation")
        # new scenarios on the same handle
        par2 = par.copy()
        par2[:, 3] = 2.0
        d.set_params(par2)
        o = oracle_mod.OracleModel("goddard", "midpoint", N, params=par2[0])
        assert_close(d.evaluate(ctd.CONS, x)["cons"][0], o.cons(x[0]), scale=max(1.0, np.abs(x[0]).max()), what="set_params")


def test_device_tensors_and_host_arrays_agree(ctd):
    import torch
    d = ctd.DOCP("goddard_all", 500, "trapeze")
    x, y = synthetic_point(d, seed=3)
    rh = d.evaluate(ctd.CONS | ctd.JAC | ctd.HESS | ctd.GRAD | ctd.OBJ, x, y, 0.3)
    xd, yd = torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda()
    out = {"jac": torch.empty(d.nnzj, dtype=torch.float64, device="cuda")}
    rd = d.evaluate(ctd.CONS | ctd.JAC | ctd.HESS | ctd.GRAD | ctd.OBJ, xd, yd, 0.3, out=out)
    d.synchronize()
    assert rd["jac"].data_ptr() == out["jac"].data_ptr()          # caller-owned buffers are filled in place
    for k in rh:
        assert np.array_equal(rh[k], rd[k].cpu().numpy()), k
    with pytest.raises(TypeError):
        d.evaluate(ctd.HESS, xd, y)                                # mixed host / device arguments
    with pytest.raises(ValueError):
        d.evaluate(ctd.CONS, x[:-1])


def test_nonuniform_time_grid(ctd, oracle_mod):
    rng = np.random.default_rng(2)
    grid = np.sort(np.concatenate(([0.0, 1.0], rng.uniform(0, 1, 200))))
    for problem, scheme in (("goddard", "midpoint"), ("bolza_freetf", "gauss_legendre_2"), ("vanderpol", "trapeze")):
        o = oracle_mod.OracleModel(problem, scheme, time_grid=grid)
        d = ctd.DOCP(problem, scheme=scheme, time_grid=grid)
        x, y = synthetic_point(o, seed=9)
        _compare_all(ctd, o, d, x, y, 0.9, f"{problem}/{scheme}/nonuniform")


def test_stagewise_known_answer_on_gpu(ctd):
    """test/ci/test_discretization_stagewise.jl:93-99 through the CUDA path: c == bounds, objective 4/3."""
    grid = np.array([0.0, 0.2, 0.6, 1.0])
    for scheme, stage, block in (("gauss_legendre_2", 2, 5), ("gauss_legendre_3", 3, 7)):
        d = ctd.DOCP("stagewise_scalar", scheme=scheme, time_grid=grid)
        nodes = {2: [0.5 - np.sqrt(3) / 6, 0.5 + np.sqrt(3) / 6], 3: [0.5 - 0.1 * np.sqrt(15), 0.5, 0.5 + 0.1 * np.sqrt(15)]}[stage]
        xu = np.zeros(d.nvar)
        for i in range(4):
            xu[i * block] = grid[i] ** 2
        for i in range(3):
            h = grid[i + 1] - grid[i]
            for j in range(stage):
                xu[i * block + 1 + j] = xu[i * block + 1 + stage + j] = 2 * (grid[i] + nodes[j] * h)
        nlp = d.nlp_model()
        _, _, lcon, ucon = d.bounds()
        c = nlp.cons(xu)
        np.testing.assert_allclose(c, lcon, atol=1e-12)
        np.testing.assert_allclose(c, ucon, atol=1e-12)
        assert abs(nlp.obj(xu) - 4.0 / 3.0) <= 1e-12


def test_solution_unpacking(ctd, oracle_mod):
    """getter / build_OCP_solution (src/ode/common.jl:7-104, src/DOCP_data.jl:514-633): copies are bit-exact,
    the time grid and the path multipliers (divided by the step length) too (same non-contracted formulas)."""
    rng = np.random.default_rng(8)
    for problem, scheme in (("goddard_all", "trapeze"), ("goddard_all", "midpoint"), ("goddard", "gauss_legendre_3"),
                            ("goddard_all", "gauss_legendre_2"), ("goddard", "euler_implicit"), ("goddard_all", "euler"),
                            ("estimate_rotation_rate", "midpoint"), ("double_integrator_minenergy", "gauss_legendre_2_constant_control")):
        o = oracle_mod.OracleModel(problem, scheme, 37)
        d = ctd.DOCP(problem, 37, scheme)
        x = o.x0() * (1 + 1e-2 * rng.uniform(-1, 1, o.nvar))
        y = rng.standard_normal(o.ncon)
        ref, got = o.unpack(x, y), d.unpack(x, y)
        assert set(got) | {"control_time"} == set(ref)
        for k in got:
            assert got[k].shape == ref[k].shape, (problem, scheme, k)
            assert np.array_equal(got[k], ref[k]), (problem, scheme, k)
        assert np.array_equal(d.control_time(got["time"]), ref["control_time"])   # = time for collocation
        prim = d.unpack(x)                       # box multipliers use the same call without y
        assert set(prim) == {"time", "state", "control", "variable"}
    # batched, device tensors
    import torch
    B = 3
    d = ctd.DOCP("goddard_all", 50, "midpoint", batch=B)
    o = oracle_mod.OracleModel("goddard_all", "midpoint", 50)
    x, y = synthetic_point(d, seed=5, batch=B)
    got = d.unpack(torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda())
    d.synchronize()
    for b in range(B):
        ref = o.unpack(x[b], y[b])
        for k in got:
            assert np.array_equal(got[k][b].cpu().numpy(), ref[k]), (b, k)


def test_every_output_entry_is_written_and_nothing_else(ctd):
    """Own bounds check (compute-sanitizer is not available on the GPU pool): outputs live inside guarded
    buffers pre-filled with a sentinel; after an evaluation no sentinel is left inside and the guards are intact."""
    import torch
    SENT = -7.25e300
    G = 64
    cases = [("goddard", "midpoint", 1000, 1, "reference"), ("goddard", "midpoint", 1000, 7, "reference"),
             ("goddard", "midpoint", 33, 5, "optimized"), ("goddard_all", "trapeze", 311, 1, "reference"),
             ("goddard", "gauss_legendre_3", 95, 1, "reference"), ("goddard", "gauss_legendre_3", 2100, 1, "optimized"),
             ("robot", "trapeze", 64, 2, "reference"), ("double_integrator_minenergy", "euler", 3, 1, "reference"),
             ("goddard", "midpoint", 2, 1, "reference"), ("goddard", "midpoint", 1, 3, "reference")]
    for p, s, N, B, m in cases:
        d = ctd.DOCP(p, N, s, batch=B, pattern_mode=m)
        x, y = synthetic_point(d, seed=21, batch=B)
        xd = torch.from_numpy(np.ascontiguousarray(x)).cuda()
        yd = torch.from_numpy(np.ascontiguousarray(y)).cuda()
        sizes = {"obj": B, "grad": B * d.nvar, "cons": B * d.ncon, "jac": B * d.nnzj, "hess": B * d.nnzh}
        for F in (ctd.CONS, ctd.JAC, ctd.HESS, ctd.GRAD | ctd.OBJ, ctd.CONS | ctd.JAC | ctd.HESS,
                  ctd.CONS | ctd.JAC | ctd.HESS | ctd.GRAD | ctd.OBJ):
            keys = [k for bit, k in ((ctd.OBJ, "obj"), (ctd.GRAD, "grad"), (ctd.CONS, "cons"), (ctd.JAC, "jac"), (ctd.HESS, "hess")) if F & bit]
            big = {k: torch.full((sizes[k] + 2 * G,), SENT, dtype=torch.float64, device="cuda") for k in keys}
            out = {k: big[k][G:G + sizes[k]].view((B, -1) if (B > 1 and k != "obj") else (-1,)) for k in keys}
            d.evaluate(F, xd, yd, 0.9, out=out)
            d.synchronize()
            for k in keys:
                b = big[k].cpu().numpy()
                tag = f"{p}/{s}/{m}/N={N}/B={B}/F={F}/{k}"
                assert np.all(b[:G] == SENT) and np.all(b[G + sizes[k]:] == SENT), tag + ": wrote outside the array"
                assert not np.any(b[G:G + sizes[k]] == SENT), tag + ": entries left unwritten"
                assert np.all(np.isfinite(b[G:G + sizes[k]])), tag


def test_guard_bands_every_compiled_model(ctd):
    """The same own bounds check for EVERY compiled (problem, scheme, pattern mode) triple on a grid of several tiles,
    fused evaluation: nothing outside the arrays, nothing left unwritten.  (A v column without per-step entries made
    the tile epilogue store one double before its range: only an exhaustive sweep over the patterns finds that.)"""
    import torch
    SENT = -7.25e300
    G = 64
    ALL = ctd.CONS | ctd.JAC | ctd.HESS | ctd.GRAD | ctd.OBJ
    for N, B in ((97, 1), (64, 3)):
        for (p, s, m) in _models(ctd):
            d = ctd.DOCP(p, N, s, batch=B, pattern_mode=m)
            x, y = synthetic_point(d, seed=3, batch=B)
            xd = torch.from_numpy(np.ascontiguousarray(x)).cuda()
            yd = torch.from_numpy(np.ascontiguousarray(y)).cuda()
            sizes = {"obj": B, "grad": B * d.nvar, "cons": B * d.ncon, "jac": B * d.nnzj, "hess": B * d.nnzh}
            big = {k: torch.full((n + 2 * G,), SENT, dtype=torch.float64, device="cuda") for k, n in sizes.items()}
            out = {k: big[k][G:G + n].view((B, -1) if (B > 1 and k != "obj") else (-1,)) for k, n in sizes.items()}
            d.evaluate(ALL, xd, yd, 0.9, out=out)
            d.synchronize()
            for k, n in sizes.items():
                b = big[k].cpu().numpy()
                tag = f"{p}/{s}/{m}/N={N}/B={B}/{k}"
                assert np.all(b[:G] == SENT) and np.all(b[G + n:] == SENT), tag + ": wrote outside the array"
                assert not np.any(b[G:G + n] == SENT), tag + ": entries left unwritten"


def test_initial_guess_packing_roundtrip(ctd, oracle_mod):
    """__initial_guess from sampled data (src/DOCP_variables.jl:122-145, irk_stagewise.jl:302-335): bit-exact
    against the oracle's restatement of the setters; pack(unpack(x)) reproduces the node variables."""
    rng = np.random.default_rng(13)
    for problem, scheme in (("goddard_all", "trapeze"), ("goddard", "midpoint"), ("goddard", "gauss_legendre_3"),
                            ("goddard", "gauss_legendre_2_constant_control"), ("estimate_rotation_rate", "euler"),
                            ("double_integrator_minenergy", "gauss_legendre_2")):
        N = 23
        o = oracle_mod.OracleModel(problem, scheme, N)
        d = ctd.DOCP(problem, N, scheme)
        st = rng.standard_normal((N + 1, o.nx))
        ct = rng.standard_normal((N + 1, o.nu)) if o.nu else None
        va = rng.standard_normal(o.nv) if o.nv else None
        assert np.array_equal(d.pack(st, ct, va), o.pack(st, ct, va)), (problem, scheme)
        assert np.array_equal(d.pack(), np.full(o.nvar, 0.1))                     # default guess 0.1 everywhere
        assert np.array_equal(d.pack(variable=va), o.pack(variable=va))
        if d.stage and "constant" not in scheme and o.nu:
            sc = rng.standard_normal((N, d.stage, o.nu))
            assert np.array_equal(d.pack(st, None, va, sc), o.pack(st, None, va, sc)), (problem, scheme)
        x = d.pack(st, ct, va)
        u = d.unpack(x)
        assert np.array_equal(u["state"], st)
        if va is not None:
            assert np.array_equal(u["variable"], va)
    import torch
    d = ctd.DOCP("goddard", 10, "midpoint", batch=2)
    st = torch.randn(2, 11, 3, dtype=torch.float64, device="cuda")
    x = d.pack(state=st)
    d.synchronize()
    assert x.shape == (2, d.nvar) and torch.equal(x[:, 0:3], st[:, 0]) and float(x[0, 3]) == 0.1


def test_wide_elements_on_large_grids_and_batches(ctd, oracle_mod):
    """Wide elements use the parts kernel on small grids; large grids and batches go through the one-warp-per-
    tile kernel with the full register file.  Both paths against the oracle."""
    o = oracle_mod.OracleModel("goddard", "gauss_legendre_2", 120_000)
    d = ctd.DOCP("goddard", 120_000, "gauss_legendre_2")
    x, y = synthetic_point(o, seed=31)
    _compare_all(ctd, o, d, x, y, 0.8, "goddard/gauss_legendre_2/N=120000")
    B, N = 60, 400
    rng = np.random.default_rng(3)
    par = np.tile(np.array([310.0, 500.0, 2.0, 3.5]), (B, 1))
    par[:, 3] = rng.uniform(1.0, 3.5, B)
    d = ctd.DOCP("goddard", N, "gauss_legendre_3_constant_control", params=par, batch=B)
    x, y = synthetic_point(d, seed=32, batch=B)
    r = d.evaluate(ctd.CONS | ctd.JAC | ctd.HESS | ctd.GRAD | ctd.OBJ, x, y, 1.0)
    obj, viol = d.norms(x)
    _, _, lcon, ucon = d.bounds()
    for b in (0, 17, 59):
        o = oracle_mod.OracleModel("goddard", "gauss_legendre_3_constant_control", N, params=par[b])
        g, j, h = o.derivs(x[b], y[b], 1.0)
        c = o.cons(x[b])
        assert_close(r["cons"][b], c, scale=max(1.0, np.abs(x[b]).max()), what=f"b={b} cons")
        assert_close(r["jac"][b], j, scale=column_scale(d.jacobian_structure(), j, d.nvar), what=f"b={b} jac")
        assert_close(r["hess"][b], h, scale=hess_scale(o, x[b], y[b], 1.0, h), what=f"b={b} hess")
        assert_close(viol[b], max(0.0, float(np.max(np.maximum(lcon - c, c - ucon)))), scale=max(1.0, np.abs(x[b]).max()),
                     what=f"b={b} violation (stand-alone norm kernel for wide elements)")


@pytest.mark.gpu
def test_native_norm_gather_single_rank(ctd):
    """The library's own NCCL communicator (ctd_comm_init / ctd_gather_norms_dev / ctd_comm_wait) with one
    rank: the gathered (objective, violation) rows of every pipelined evaluation equal ctd_norms."""
    import torch
    from ctdirect_b200.batch import NativeNormsPipeline
    B, N = 6, 200
    rng = np.random.default_rng(5)
    par = np.tile(np.array([310.0, 500.0, 2.0, 3.5]), (B, 1))
    par[:, 3] = rng.uniform(1.0, 3.5, B)
    d = ctd.DOCP("goddard", N, "midpoint", params=par, batch=B)
    assert ctd._lib.load().ctd_nccl_version() > 20000
    pipe = NativeNormsPipeline(d, 0, 1, ctd.DOCP.nccl_unique_id(), depth=2)
    dev = torch.device("cuda", 0)
    pts = []
    for k in range(5):
        x, y = synthetic_point(d, seed=100 + k, batch=B)
        pts.append((torch.from_numpy(x).to(dev), torch.from_numpy(y).to(dev)))
    ALL = ctd.CONS | ctd.JAC | ctd.HESS | ctd.GRAD | ctd.OBJ
    for k, (tx, ty) in enumerate(pts):
        s = pipe.submit(ALL, tx, ty, serial=(k == 0))
        g = pipe.result(s)
        d.synchronize()
        obj, viol = d.norms(tx)
        assert g.shape == (1, 2, B)
        assert_close(g[0, 0].cpu().numpy(), obj.cpu().numpy(), scale=1.0, what=f"evaluation {k}: gathered objective")
        assert_close(g[0, 1].cpu().numpy(), viol.cpu().numpy(), scale=1.0, what=f"evaluation {k}: gathered violation")
    pipe.drain()
    d.synchronize()


def _native_gather_worker(rank, world, port, q, kind="native", rounds=3):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)   # only carries the 128-byte NCCL id
    try:
        import ctdirect_b200 as ctd
        from ctdirect_b200.batch import NativeNormsPipeline, PeerNormsPipeline, broadcast_unique_id, shard_range
        B, N = 8, 120
        rng = np.random.default_rng(3)
        par = np.tile(np.array([310.0, 500.0, 2.0, 3.5]), (B, 1))
        par[:, 3] = rng.uniform(1.0, 3.5, B)
        lo, hi = shard_range(B, world, rank)
        d = ctd.DOCP("goddard", N, "midpoint", params=par[lo:hi], batch=hi - lo, device=rank)
        dev = torch.device("cuda", rank)
        x_all = np.tile(d.initial_guess(), (B, 1)) * (1 + 1e-3 * rng.uniform(-1, 1, (B, d.nvar)))
        res = []
        if kind == "split":   # the shard as two half-shards on two streams, each with its own fused gather
            from ctdirect_b200.batch import SplitShardPipeline
            mid = (lo + hi) // 2
            parts = [ctd.DOCP("goddard", N, "midpoint", params=par[a:b], batch=b - a, device=rank, stream="own")
                     for a, b in ((lo, mid), (mid, hi))]
            split = SplitShardPipeline(parts, rank, world, depth=4)
            for k in range(rounds):
                xs = [torch.from_numpy(np.ascontiguousarray(x_all[a:b] * (1 + 1e-4 * k))).to(dev) for a, b in ((lo, mid), (mid, hi))]
                torch.cuda.synchronize()
                sl = split.submit(ctd.CONS | ctd.OBJ, xs)
                gs = split.result(sl)
                for p_ in parts:
                    p_.synchronize()
                # (world, 2, half) per part -> (world, 2, shard) in instance order
                res.append(np.concatenate([g.cpu().numpy() for g in gs], axis=2))
            split.drain()
            q.put((rank, res))
            return
        pipe = (PeerNormsPipeline if kind == "peer" else NativeNormsPipeline)(d, rank, world, broadcast_unique_id(), depth=2)
        for k in range(rounds):
            xk = torch.from_numpy(np.ascontiguousarray(x_all[lo:hi] * (1 + 1e-4 * k))).to(dev)
            s = pipe.submit(ctd.CONS | ctd.OBJ, xk)
            g = pipe.result(s)
            d.synchronize()
            res.append(g.cpu().numpy().copy())
        pipe.drain()
        d.synchronize()
        q.put((rank, res))
    finally:
        dist.destroy_process_group()


@pytest.mark.gpu
def test_native_norm_gather_two_ranks(ctd):
    """Two processes, two GPUs: every rank ends up with the (objective, violation) rows of both shards, equal to
    a single-process evaluation of the whole batch.  Skipped on a one-GPU box."""
    import socket
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mpc = mp.get_context("spawn")
    q = mpc.Queue()
    procs = [mpc.Process(target=_native_gather_worker, args=(r, 2, port, q)) for r in range(2)]
    for p_ in procs:
        p_.start()
    out = dict(q.get(timeout=180) for _ in range(2))
    for p_ in procs:
        p_.join(timeout=60)
        assert p_.exitcode == 0
    B, N = 8, 120
    rng = np.random.default_rng(3)
    par = np.tile(np.array([310.0, 500.0, 2.0, 3.5]), (B, 1))
    par[:, 3] = rng.uniform(1.0, 3.5, B)
    d = ctd.DOCP("goddard", N, "midpoint", params=par, batch=B)
    x_all = np.tile(d.initial_guess(), (B, 1)) * (1 + 1e-3 * rng.uniform(-1, 1, (B, d.nvar)))
    for k in range(3):
        obj, viol = d.norms(np.ascontiguousarray(x_all * (1 + 1e-4 * k)))
        for rank in (0, 1):
            g = out[rank][k]                    # (world, 2, 4)
            assert_close(g[:, 0, :].reshape(-1), obj, scale=1.0, what=f"rank {rank} evaluation {k}: objectives")
            assert_close(g[:, 1, :].reshape(-1), viol, scale=1.0, what=f"rank {rank} evaluation {k}: violations")


@pytest.mark.gpu
def test_peer_memory_gather_two_ranks(ctd):
    """The gather fused into the evaluation kernel (NVLink peer stores from the reduction epilogue, ctd_eval_gather_dev):
    over 7 rounds on 2 slots (every slot recycled three times) each rank holds the rows of both shards, bit-identical
    to a single-process evaluation of the whole batch.  Skipped on a one-GPU box."""
    import socket
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mpc = mp.get_context("spawn")
    q = mpc.Queue()
    R = 7
    procs = [mpc.Process(target=_native_gather_worker, args=(r, 2, port, q, "peer", R)) for r in range(2)]
    for p_ in procs:
        p_.start()
    out = dict(q.get(timeout=180) for _ in range(2))
    for p_ in procs:
        p_.join(timeout=60)
        assert p_.exitcode == 0
    B, N = 8, 120
    rng = np.random.default_rng(3)
    par = np.tile(np.array([310.0, 500.0, 2.0, 3.5]), (B, 1))
    par[:, 3] = rng.uniform(1.0, 3.5, B)
    d = ctd.DOCP("goddard", N, "midpoint", params=par, batch=B)
    x_all = np.tile(d.initial_guess(), (B, 1)) * (1 + 1e-3 * rng.uniform(-1, 1, (B, d.nvar)))
    for k in range(R):
        obj, viol = d.norms(np.ascontiguousarray(x_all * (1 + 1e-4 * k)))
        for rank in (0, 1):
            g = out[rank][k]                    # (world, 2, 4)
            assert np.array_equal(g[:, 0, :].reshape(-1), obj), f"rank {rank} round {k}: objectives"
            assert np.array_equal(g[:, 1, :].reshape(-1), viol), f"rank {rank} round {k}: violations"


@pytest.mark.gpu
def test_split_shard_pipeline_two_ranks(ctd):
    """batch.SplitShardPipeline: every rank's shard as two half-shards on two streams, each with its own communicator
    and fused peer-memory gather (the best 8-GPU schedule of bench.py).  Over 9 rounds on 4 slots the rows of both
    halves of both ranks are bit-identical to a single-process evaluation of the whole batch."""
    import socket
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mpc = mp.get_context("spawn")
    q = mpc.Queue()
    R = 9
    procs = [mpc.Process(target=_native_gather_worker, args=(r, 2, port, q, "split", R)) for r in range(2)]
    for p_ in procs:
        p_.start()
    out = dict(q.get(timeout=180) for _ in range(2))
    for p_ in procs:
        p_.join(timeout=60)
        assert p_.exitcode == 0
    B, N = 8, 120
    rng = np.random.default_rng(3)
    par = np.tile(np.array([310.0, 500.0, 2.0, 3.5]), (B, 1))
    par[:, 3] = rng.uniform(1.0, 3.5, B)
    d = ctd.DOCP("goddard", N, "midpoint", params=par, batch=B)
    x_all = np.tile(d.initial_guess(), (B, 1)) * (1 + 1e-3 * rng.uniform(-1, 1, (B, d.nvar)))
    for k in range(R):
        obj, viol = d.norms(np.ascontiguousarray(x_all * (1 + 1e-4 * k)))
        for rank in (0, 1):
            g = out[rank][k]                    # (world, 2, 4): rank-major, each rank's two halves side by side
            assert np.array_equal(g[:, 0, :].reshape(-1), obj), f"rank {rank} round {k}: objectives"
            assert np.array_equal(g[:, 1, :].reshape(-1), viol), f"rank {rank} round {k}: violations"


@pytest.mark.gpu
def test_caller_stream_and_graph_capture(ctd):
    """ctd_set_stream: the evaluation runs in the caller's stream order and can be captured into a CUDA graph;
    replays reproduce the eager results bit for bit on new inputs written into the same buffers."""
    import torch
    dev = torch.device("cuda", 0)
    d = ctd.DOCP("goddard", 3000, "midpoint")
    F = ctd.CONS | ctd.JAC | ctd.HESS
    x0, y0 = synthetic_point(d, seed=1)
    x1, y1 = synthetic_point(d, seed=2)
    ref0 = d.evaluate(F, torch.from_numpy(x0).to(dev), torch.from_numpy(y0).to(dev), 1.0)
    ref1 = d.evaluate(F, torch.from_numpy(x1).to(dev), torch.from_numpy(y1).to(dev), 1.0)
    d.synchronize()
    st = torch.cuda.Stream(device=dev)
    d.set_stream(st)
    assert d.cuda_stream == st.cuda_stream
    x, y = torch.empty(d.nvar, dtype=torch.float64, device=dev), torch.empty(d.ncon, dtype=torch.float64, device=dev)
    call, out = d.bind(F, x, y, 1.0)
    with torch.cuda.stream(st):
        x.copy_(torch.from_numpy(x0).to(dev)); y.copy_(torch.from_numpy(y0).to(dev))
        call()                                  # ordered after the copies by the caller's stream
    st.synchronize()
    for k in ("cons", "jac", "hess"):
        assert torch.equal(out[k], ref0[k]), k
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=st, capture_error_mode="thread_local"):
        call()
    for xs, ys, ref in ((x1, y1, ref1), (x0, y0, ref0)):
        with torch.cuda.stream(st):
            x.copy_(torch.from_numpy(xs).to(dev)); y.copy_(torch.from_numpy(ys).to(dev))
            for v in out.values():
                v.zero_()
            g.replay()
        st.synchronize()
        for k in ("cons", "jac", "hess"):
            assert torch.equal(out[k], ref[k]), k
    d.set_stream(None)
    d.evaluate(F, x, y, 1.0)
    d.synchronize()


@pytest.mark.gpu
@pytest.mark.parametrize("problem,scheme,N,mode,batch", [
    ("goddard", "midpoint", 500, "reference", 1), ("goddard", "midpoint", 500, "optimized", 3),
    ("goddard", "gauss_legendre_3", 60, "reference", 1), ("robot", "trapeze", 50, "optimized", 2),
    ("goddard", "midpoint", 100_000, "reference", 1),
])
def test_kkt_assembly(ctd, problem, scheme, N, mode, batch):
    """ctd_kkt_assemble_dev: the values of one evaluation moved to the CSC order of
    tril([[H + diag(dx), J'], [J, -diag(dc)]]) equal scipy's assembly of the same arrays, exactly (the only
    arithmetic is H_jj + dx_j)."""
    import torch
    from test_host_structure import _kkt_reference
    dev = torch.device("cuda", 0)
    d = ctd.DOCP(problem, N, scheme, pattern_mode=mode, batch=batch)
    x, y = synthetic_point(d, seed=21, batch=batch)
    r = d.evaluate(ctd.CONS | ctd.JAC | ctd.HESS, torch.from_numpy(x).to(dev), torch.from_numpy(y).to(dev), 1.0)
    rng = np.random.default_rng(4)
    dx, dc = rng.uniform(0.1, 2.0, (batch, d.nvar)), rng.uniform(0.1, 2.0, (batch, d.ncon))
    k = d.kkt_assemble(r["hess"], r["jac"], torch.from_numpy(dx).to(dev), torch.from_numpy(dc).to(dev))
    k0 = d.kkt_assemble(r["hess"], r["jac"])     # no regularisation
    d.synchronize()
    k, k0 = k.cpu().numpy().reshape(batch, -1), k0.cpu().numpy().reshape(batch, -1)
    hv, jv = r["hess"].cpu().numpy().reshape(batch, -1), r["jac"].cpu().numpy().reshape(batch, -1)
    colptr, rows = d.kkt_structure()
    for b in range(batch):
        K = _kkt_reference(d, hv[b], jv[b], dx[b], dc[b])
        assert np.array_equal(colptr - 1, K.indptr) and np.array_equal(rows - 1, K.indices)
        assert np.array_equal(k[b], K.data), f"instance {b}"
        assert np.array_equal(k0[b], _kkt_reference(d, hv[b], jv[b]).data), f"instance {b}, no regularisation"


@pytest.mark.gpu
def test_nan_and_inf_pass_through(ctd):
    """Like the reference, a NaN / Inf iterate is not an error: it flows into the results the solver reads
    (SURVEY.md 8b).  Only the entries that depend on the poisoned step change; the violation norm becomes NaN."""
    import torch
    dev = torch.device("cuda", 0)
    N = 400
    d = ctd.DOCP("goddard", N, "midpoint")
    x, y = synthetic_point(d, seed=8)
    F = ctd.CONS | ctd.JAC | ctd.HESS
    clean = d.evaluate(F, x, y, 1.0)
    step = 200                                      # 1-based step whose state is poisoned
    for bad in (np.nan, np.inf):
        xb = x.copy()
        xb[(step - 1) * 4 + 1] = bad                # v_200
        r = d.evaluate(F, xb, y, 1.0)
        c = r["cons"].reshape(-1)
        rows = slice((step - 2) * 3, step * 3)      # defects of steps 199 and 200 read x_200
        assert not np.all(np.isfinite(c[rows]))
        far = np.ones(d.ncon, dtype=bool)
        far[rows] = False
        assert np.array_equal(c[far], clean["cons"].reshape(-1)[far])
        jr, jc = d.jacobian_structure()
        touched = np.isin((jr - 1) // 3, (step - 2, step - 1)) & (jr <= 3 * N)
        assert np.array_equal(r["jac"].reshape(-1)[~touched], clean["jac"].reshape(-1)[~touched])
        assert not np.all(np.isfinite(r["jac"].reshape(-1)[touched]))
        assert not np.all(np.isfinite(r["hess"]))
        obj, viol = d.norms(xb)
        assert np.isnan(viol[0]) or np.isinf(viol[0])
    tx = torch.from_numpy(x).to(dev)
    assert np.isfinite(d.norms(tx)[1].cpu().numpy()).all()


@pytest.mark.gpu
def test_cuda_path_against_committed_regression_vectors(ctd):
    """The CUDA path against tests/golden/regression_vectors.npz (oracle-made regression anchors, see
    tests/golden/make_regression_vectors.py): same tolerance as the live comparisons."""
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "regression_vectors.npz"))
    tags = sorted({k.rsplit("|", 1)[0] for k in z.files})
    for tag in tags:
        p, s, n, mode = tag.split("|")
        d = ctd.DOCP(p, int(n), s, pattern_mode=mode)
        x, y = z[tag + "|x"], z[tag + "|y"]
        r = d.evaluate(ctd.CONS | ctd.JAC | ctd.HESS | ctd.GRAD | ctd.OBJ, x, y, 0.75)
        xs = max(1.0, float(np.abs(x).max()))
        assert_close(r["cons"], z[tag + "|cons"], scale=xs, what=tag + " cons")
        assert_close(r["jac"], z[tag + "|jac"], scale=column_scale(d.jacobian_structure(), z[tag + "|jac"], d.nvar), what=tag + " jac")
        from oracle import oracle as orc   # (test infrastructure: the tolerance floor of the Hessian needs the per-row terms)
        o = orc.OracleModel(p, s, int(n), pattern_mode=mode)
        assert_close(r["hess"], z[tag + "|hess"], scale=hess_scale(o, x, y, 0.75, z[tag + "|hess"]), what=tag + " hess")
        assert_close(r["grad"], z[tag + "|grad"], scale=1e-3, what=tag + " grad")
        assert_close(r["obj"], z[tag + "|obj"], scale=1.0, what=tag + " obj")


@pytest.mark.gpu
def test_library_call_counters(ctd):
    """The per-handle counters of the C library (ctd_counters / ctd_stats_json) follow NLPModels.Counters: every
    entry point counts the outputs it produced, fused or not (reference: the counters printed by test/benchmark.jl:81-125)."""
    import torch
    d = ctd.DOCP("goddard", 40, "midpoint")
    nlp = d.nlp_model()
    x, y = synthetic_point(d, seed=5)
    assert d.counters() == dict.fromkeys(d.COUNTER_NAMES, 0)
    nlp.obj(x); nlp.grad(x); nlp.cons(x); nlp.cons(x); nlp.jac_coord(x); nlp.hess_coord(x, y)
    nlp.eval_all(x, y)
    xd, yd = torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda()
    d.evaluate(ctd.CONS | ctd.JAC | ctd.HESS, xd, yd)
    d.synchronize()
    c = d.counters()
    # grad evaluates the objective along with the gradient (ctd_grad = CTD_GRAD | CTD_OBJ)
    assert c == {"neval_obj": 3, "neval_grad": 2, "neval_cons": 4, "neval_jac": 3, "neval_hess": 3,
                 "calls": 8, "host_calls": 7, "norm_calls": 0}, c
    s = d.stats()
    assert s["problem"] == "goddard" and s["scheme"] == "midpoint" and s["grid_size"] == 40
    assert (s["nvar"], s["ncon"], s["nnzj"], s["nnzh"]) == (d.nvar, d.ncon, d.nnzj, d.nnzh)
    assert s["neval_cons"] == 4 and s["kernel_launches"] == d.launch_count >= 8 and s["host_call_seconds"] > 0
    d.reset_counters()
    assert d.counters() == dict.fromkeys(d.COUNTER_NAMES, 0)


@pytest.mark.gpu
def test_implicit_euler_with_path_constraints(ctd, oracle_mod):
    """goddard_all / euler_implicit: node i reads the control of block i-1 (src/ode/euler.jl:66-69), so the step
    elements carry the path rows of the NEXT node, node 1 is an extra job of the edge CTA and node N+1 the final-path
    element.  Only the corrected pattern holds these dependencies (the reference's :manual pattern lists U_i,
    euler.jl:214-216).  Values against the oracle at sizes around the tile boundaries, batches, the in-kernel
    violation norm, solution unpacking."""
    with pytest.raises(ctd.CtdError, match="not compiled in"):
        ctd.DOCP("goddard_all", 10, "euler_implicit")                       # reference pattern: refused
    lv = None
    for N in (1, 2, 3, 4, 31, 32, 33, 34, 63, 64, 65, 1000):
        # the detected pattern (the reference's default backend) holds the dependency by construction
        od = oracle_mod.OracleModel("goddard_all", "euler_implicit", N, pattern_mode="optimized")
        dd = ctd.DOCP("goddard_all", N, "euler_implicit", pattern_mode="optimized")
        for a, b in zip(dd.jacobian_structure() + dd.hessian_structure(), od.jac_structure() + od.hess_structure()):
            assert np.array_equal(a, b), N
        x, y = synthetic_point(od, seed=300 + N)
        _compare_all(ctd, od, dd, x, y, 0.6, f"goddard_all/euler_implicit/optimized/N={N}")
        o = oracle_mod.OracleModel("goddard_all", "euler_implicit", N, pattern_mode="corrected")
        d = ctd.DOCP("goddard_all", N, "euler_implicit", pattern_mode="corrected")
        _compare_all(ctd, o, d, x, y, 0.6, f"goddard_all/euler_implicit/corrected/N={N}")
        # the Jacobian holds the dependency of node i's path rows on the control of block i-1
        if N >= 3:
            r, c = d.jacobian_structure()
            B, CB = d.step_variables_block, d.step_constraints_block
            rows_u1 = set(r[c == d.NLP_x + 1].tolist())                      # column of U_1
            assert {d.NLP_x + q + 1 for q in range(d.path_cons)} <= rows_u1          # node 1 (u(t_1) = U_1)
            assert {CB + d.NLP_x + q + 1 for q in range(d.path_cons)} <= rows_u1     # node 2
            rows_u2 = set(r[c == B + d.NLP_x + 1].tolist())
            assert not ({CB + d.NLP_x + q + 1 for q in range(d.path_cons)} & rows_u2)
        obj, viol = d.norms(x)
        lvar, uvar, lcon, ucon = d.bounds()
        cref = o.cons(x)
        v_ref = max(0.0, float(np.max(np.maximum(lcon - cref, cref - ucon))))
        assert_close(viol[0], v_ref, scale=max(1.0, np.abs(x).max()), what=f"N={N} violation")
        u_ref, u_got = o.unpack(x, y), d.unpack(x, y)
        for k in u_got:
            assert np.array_equal(u_got[k], u_ref[k]), (N, k)
    # a non-uniform time grid (the step lengths enter the shifted rows through t_{i+1})
    rng = np.random.default_rng(4)
    grid = np.concatenate([[0.0], np.sort(rng.uniform(0.02, 0.98, 69)), [1.0]])
    for mode in ("corrected", "optimized"):
        o = oracle_mod.OracleModel("goddard_all", "euler_implicit", 70, time_grid=grid, pattern_mode=mode)
        d = ctd.DOCP("goddard_all", 70, "euler_implicit", time_grid=grid, pattern_mode=mode)
        x, y = synthetic_point(o, seed=12)
        _compare_all(ctd, o, d, x, y, 1.3, f"goddard_all/euler_implicit/{mode}/non-uniform grid")
    B, N = 7, 100
    d = ctd.DOCP("goddard_all", N, "euler_implicit", pattern_mode="corrected", batch=B)
    o = oracle_mod.OracleModel("goddard_all", "euler_implicit", N, pattern_mode="corrected")
    x, y = synthetic_point(d, seed=8, batch=B)
    r = d.evaluate(ctd.CONS | ctd.JAC | ctd.HESS | ctd.GRAD | ctd.OBJ, x, y, 1.0)
    obj, viol = d.norms(x)
    lvar, uvar, lcon, ucon = d.bounds()
    for b in range(B):
        g, j, h = o.derivs(x[b], y[b], 1.0)
        cref = o.cons(x[b])
        assert_close(r["cons"][b], cref, scale=max(1.0, np.abs(x[b]).max()), what=f"shifted rows, instance {b} cons")
        assert_close(r["jac"][b], j, scale=column_scale(d.jacobian_structure(), j, d.nvar), what=f"shifted rows, instance {b} jac")
        assert_close(r["hess"][b], h, scale=hess_scale(o, x[b], y[b], 1.0, h), what=f"shifted rows, instance {b} hess")
        assert_close(viol[b], max(0.0, float(np.max(np.maximum(lcon - cref, cref - ucon)))),
                     scale=max(1.0, np.abs(x[b]).max()), what=f"shifted rows, instance {b} violation")
