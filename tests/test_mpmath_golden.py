"""Full Jacobian / Lagrangian-Hessian entries of the discretised Goddard problems at N = 3 against a 50-digit evaluation
that shares nothing with the oracle, the generator or the kernels (tests/golden/make_mpmath_golden.py: the NLP functions
written out from the reference's formulas as sympy expressions, differentiated symbolically, evaluated with mpmath).
The oracle is checked here on the CPU, the CUDA path on the GPU; entries of the `:manual` pattern that the golden
derivative does not have are structural zeros and must be exactly 0."""
import os

import numpy as np
import pytest

from conftest import assert_close, column_scale, hess_scale

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mpmath_goddard_n3.npz"))
# (problem, scheme, pattern mode, key in the golden file)
CASES = [("goddard", "midpoint", "reference", "midpoint"),
         ("goddard", "gauss_legendre_3_constant_control", "reference", "gauss_legendre_3_constant_control"),
         # implicit Euler with path constraints: node i reads the control of block i-1 (euler.jl:66-69)
         ("goddard_all", "euler_implicit", "corrected", "goddard_all|euler_implicit"),
         ("goddard_all", "euler_implicit", "optimized", "goddard_all|euler_implicit"),
         # two controls per time step (DirectShooting, midpoint.jl:145-154)
         ("goddard", "midpoint_cs2", "optimized", "midpoint_cs2")]
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


@pytest.mark.parametrize("problem,scheme,mode,key", CASES)
def test_oracle_against_50_digit_entries(oracle_mod, problem, scheme, mode, key):
    o = oracle_mod.OracleModel(problem, scheme, N, pattern_mode=mode)
    x, y, sigma = GOLD[key + "|x"], GOLD[key + "|y"], float(GOLD[key + "|sigma"])
    assert GOLD[key + "|jac"].shape == (o.ncon, o.nvar)
    _, j, h = o.derivs(x, y, sigma)
    _check(f"oracle vs mpmath {problem}/{scheme}", (o.jac_structure(), o.hess_structure()),
           {"cons": o.cons(x), "jac": j, "hess": h, "obj": o.obj(x)}, o, key)
    if mode == "optimized":   # the detected pattern: the nonzeros of the 50-digit derivative plus entries that are
        J, H = GOLD[key + "|jac"], GOLD[key + "|hess"]   # structurally present but vanish (u * 0 in F0 + u F1: the tracer does not look at values)
        assert o.nnzj >= int((J != 0).sum()) and o.nnzh >= int((H != 0).sum())
        if key == "goddard_all|euler_implicit":
            assert (o.nnzj, o.nnzh) == (int((J != 0).sum()), int((H != 0).sum())) == (74, 33)


@pytest.mark.gpu
@pytest.mark.parametrize("problem,scheme,mode,key", CASES)
def test_cuda_path_against_50_digit_entries(ctd, oracle_mod, problem, scheme, mode, key):
    o = oracle_mod.OracleModel(problem, scheme, N, pattern_mode=mode)      # (only for the tolerance floor of the Hessian)
    d = ctd.DOCP(problem, N, scheme, pattern_mode=mode)
    x, y, sigma = GOLD[key + "|x"], GOLD[key + "|y"], float(GOLD[key + "|sigma"])
    r = d.evaluate(ctd.CONS | ctd.JAC | ctd.HESS | ctd.GRAD | ctd.OBJ, x, y, sigma)
    _check(f"CUDA vs mpmath {problem}/{scheme}", (d.jacobian_structure(), d.hessian_structure()),
           {"cons": r["cons"], "jac": r["jac"], "hess": r["hess"], "obj": float(r["obj"][0])}, o, key)
