"""Writes tests/golden/regression_vectors.npz: value fixtures for a few small models, produced by the CPU oracle
(oracle/, hyper-dual derivatives over the restated reference functions) at seeded points.

These are NOT reference outputs -- the reference is Julia + ADNLPModels and cannot run in this image, and it stores
no value vectors for this path ("parity unpinned" for values).  They are regression anchors: a CPU
test checks that the oracle still reproduces them, a GPU test checks the CUDA path against them, so a change in
either shows up even when both drift together in a live comparison.

    python tests/golden/make_regression_vectors.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import oracle as orc  # noqa: E402
from conftest import synthetic_point  # noqa: E402

CASES = [("goddard", "midpoint", 12, "reference"), ("goddard", "midpoint", 12, "optimized"),
         ("goddard_all", "trapeze", 9, "reference"), ("goddard", "gauss_legendre_3", 5, "reference"),
         ("goddard", "gauss_legendre_2_constant_control", 6, "reference"),
         ("double_integrator_minenergy", "trapeze", 100, "reference"), ("robot", "trapeze", 6, "reference"),
         ("vanderpol", "euler", 10, "reference"), ("bolza_freetf", "euler_implicit", 7, "reference")]


def main():
    orc.build()
    out = {}
    for k, (p, s, n, mode) in enumerate(CASES):
        o = orc.OracleModel(p, s, n, pattern_mode=mode)
        x, y = synthetic_point(o, seed=4242 + k)
        g, j, h = o.derivs(x, y, 0.75)
        tag = f"{p}|{s}|{n}|{mode}"
        out[tag + "|x"], out[tag + "|y"] = x, y
        out[tag + "|cons"], out[tag + "|jac"], out[tag + "|hess"], out[tag + "|grad"] = o.cons(x), j, h, g
        out[tag + "|obj"] = np.array([o.obj(x)])
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "regression_vectors.npz"), **out)
    print("wrote", len(CASES), "cases")


if __name__ == "__main__":
    main()
