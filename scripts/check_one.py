"""Parity of ONE (problem, scheme) of a variant library against the oracle: check_one.py problem scheme [mode]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import ctdirect_b200 as ctd
from oracle import oracle as orc
from conftest import assert_close, column_scale, hess_scale, synthetic_point
orc.build()
p, s = sys.argv[1], sys.argv[2]
mode = sys.argv[3] if len(sys.argv) > 3 else "reference"
for N in (1, 2, 3, 5, 29, 30, 31, 130, 257, 20000):
    o = orc.OracleModel(p, s, N, pattern_mode=mode)
    d = ctd.DOCP(p, N, s, pattern_mode=mode)
    x, y = synthetic_point(o, seed=100 + N)
    g_ref, j_ref, h_ref = o.derivs(x, y, 0.7)
    sh = hess_scale(o, x, y, 0.7, h_ref)
    sj = column_scale(o.jac_structure(), j_ref, o.nvar)
    for F, keys in ((31, ("cons", "jac", "hess", "grad", "obj")), (7, ("cons", "jac", "hess")), (2, ("jac",)), (4, ("hess",))):
        r = d.evaluate(F, x, y, 0.7)
        tag = f"{p}/{s}/{mode}/N={N}/F={F}"
        if "cons" in keys: assert_close(r["cons"], o.cons(x), scale=max(1.0, np.abs(x).max()), what=tag + " cons")
        if "jac" in keys: assert_close(r["jac"], j_ref, scale=sj, what=tag + " jac")
        if "hess" in keys: assert_close(r["hess"], h_ref, scale=sh, what=tag + " hess")
        if "grad" in keys: assert_close(r["grad"], g_ref, scale=1e-3 * max(1.0, np.abs(g_ref).max()), what=tag + " grad")
        if "obj" in keys: assert_close(r["obj"], o.obj(x), scale=1.0, what=tag + " obj")
        r2 = d.evaluate(F, x, y, 0.7)
        for k in keys:
            assert np.array_equal(r[k], r2[k]), tag + " not reproducible"
    print("ok", p, s, mode, N, flush=True)
