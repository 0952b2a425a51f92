"""Parity of the CUDA path (through the C ABI) with the CPU oracle.  Run on a B200: pytest -m gpu.

Tolerance (BASELINE.json north_star): indices bit-exact, values within 1e-12 relative.  `assert_close`
uses |a-b| <= 1e-12 * max(|a|, |b|, scale) where `scale` is the magnitude of the quantities that were
combined to form the entry (stated at each call)."""
import numpy as np
import pytest

from conftest import assert_close, synthetic_point

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
        sj = max(1.0, float(np.abs(j_ref).max()))
        sh = max(1e-300, float(np.abs(h_ref).max()))
        sg = max(1.0, float(np.abs(g_ref).max()))
        # separate callbacks
        assert_close(nlp.cons(x), c_ref, scale=sx, rtol=RTOL, what=tag + " cons")
        assert_close(nlp.jac_coord(x), j_ref, scale=sj * 1e-3, rtol=RTOL, what=tag + " jac")
        assert_close(nlp.hess_coord(x, y, obj_weight=sigma), h_ref, scale=sh * 1e-3, rtol=RTOL, what=tag + " hess")
        assert_close(nlp.obj(x), f_ref, scale=1.0, rtol=RTOL, what=tag + " obj")
        assert_close(nlp.grad(x), g_ref, scale=sg * 1e-3, rtol=RTOL, what=tag + " grad")
        # fused launch gives the same numbers as the separate ones
        r = nlp.eval_all(x, y, obj_weight=sigma)
        assert_close(r["cons"], c_ref, scale=sx, rtol=RTOL, what=tag + " fused cons")
        assert_close(r["jac"], j_ref, scale=sj * 1e-3, rtol=RTOL, what=tag + " fused jac")
        assert_close(r["hess"], h_ref, scale=sh * 1e-3, rtol=RTOL, what=tag + " fused hess")
        assert_close(r["grad"], g_ref, scale=sg * 1e-3, rtol=RTOL, what=tag + " fused grad")
        assert_close(r["obj"], f_ref, scale=1.0, rtol=RTOL, what=tag + " fused obj")
