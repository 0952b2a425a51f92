"""Full Jacobian / Lagrangian-Hessian entries of the discretised Goddard problem at N = 3 against a 50-digit evaluation
that shares nothing with the oracle, the generator or the kernels (tests/golden/make_mpmath_golden.py: the NLP functions
written out from the reference's formulas as sympy expressions, differentiated symbolically, evaluated with mpmath).
The oracle is checked here on the CPU, the CUDA path on the GPU; entries of the `:manual` pattern that the golden
derivative does not have are structural zeros and must be exactly 0."""
import os

import numpy as np
import pytest

from conftest import assert_close, column_scale, hess_scale

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mpmath_goddard_n3.npz"))
SCHEMES = ("midpoint", "gauss_legendre_3_constant_control")
N = 3


def _check(tag, structs, vals, o, scheme):
    (jr, jc), (hr, hc) = structs
    x, y, sigma = GOLD[scheme + "|x"], GOLD[scheme + "|y"], float(GOLD[scheme + "|sigma"])
    J, H = GOLD[scheme + "|jac"], GOLD[scheme + "|hess"]
    j_gold, h_gold = J[jr - 1, jc - 1], H[hr - 1, hc - 1]
    # the pattern covers every nonzero of the 50-digit derivative
    mj = np.zeros_like(J, dtype=bool); mj[jr - 1, jc - 1] = True
    mh = np.zeros_like(H, dtype=bool); mh[hr - 1, hc - 1] = True
    assert not (J[~mj] != 0).any() and not (H[~mh] != 0).any(), tag + ": a nonzero derivative outside the pattern"
    assert_close(vals["cons"], GOLD[scheme + "|cons"], scale=max(1.0, np.abs(x).max()), what=tag + " cons")
    assert_close(vals["jac"], j_gold, scale=column_scale((jr, jc), j_gold, o.nvar), what=tag + " jac")
    assert_close(vals["hess"], h_gold, scale=hess_scale(o, x, y, sigma, h_gold), what=tag + " hess")
    assert_close(vals["obj"], float(GOLD[scheme + "|obj"]), scale=1.0, what=tag + " obj")
    assert np.all(vals["jac"][j_gold == 0] == 0) and np.all(vals["hess"][h_gold == 0] == 0), tag + ": structural zeros"


@pytest.mark.parametrize("scheme", SCHEMES)
def test_oracle_against_50_digit_entries(oracle_mod, scheme):
    o = oracle_mod.OracleModel("goddard", scheme, N)
    x, y, sigma = GOLD[scheme + "|x"], GOLD[scheme + "|y"], float(GOLD[scheme + "|sigma"])
    assert GOLD[scheme + "|jac"].shape == (o.ncon, o.nvar)
    _, j, h = o.derivs(x, y, sigma)
    _check(f"oracle vs mpmath goddard/{scheme}", (o.jac_structure(), o.hess_structure()),
           {"cons": o.cons(x), "jac": j, "hess": h, "obj": o.obj(x)}, o, scheme)


@pytest.mark.gpu
@pytest.mark.parametrize("scheme", SCHEMES)
def test_cuda_path_against_50_digit_entries(ctd, oracle_mod, scheme):
    o = oracle_mod.OracleModel("goddard", scheme, N)      # (only for the tolerance floor of the Hessian)
    d = ctd.DOCP("goddard", N, scheme)
    x, y, sigma = GOLD[scheme + "|x"], GOLD[scheme + "|y"], float(GOLD[scheme + "|sigma"])
    r = d.evaluate(ctd.CONS | ctd.JAC | ctd.HESS | ctd.GRAD | ctd.OBJ, x, y, sigma)
    _check(f"CUDA vs mpmath goddard/{scheme}", (d.jacobian_structure(), d.hessian_structure()),
           {"cons": r["cons"], "jac": r["jac"], "hess": r["hess"], "obj": float(r["obj"][0])}, o, scheme)
