"""Small evaluation set for compute-sanitizer (one tool per run): every kernel flavour once."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ctdirect_b200 as ctd
rng = np.random.default_rng(0)
cases = [("goddard", "midpoint", 70, 1, "reference"), ("goddard", "midpoint", 40, 3, "reference"),
         ("goddard_all", "trapeze", 70, 1, "reference"), ("goddard", "gauss_legendre_3", 40, 1, "reference"),
         ("goddard", "gauss_legendre_3", 700, 1, "reference"), ("robot", "trapeze", 70, 1, "optimized"),
         ("double_integrator_minenergy", "euler_implicit", 3, 1, "reference"), ("goddard", "midpoint", 2, 1, "optimized")]
for p, s, N, B, m in cases:
    d = ctd.DOCP(p, N, s, batch=B, pattern_mode=m)
    x = np.tile(d.initial_guess(), (B, 1)) * (1 + 1e-3 * rng.uniform(-1, 1, (B, d.nvar)))
    y = rng.standard_normal((B, d.ncon))
    x, y = (x[0], y[0]) if B == 1 else (x, y)
    for F in (1, 2, 4, 16, 24, 7, 31):
        d.evaluate(F, x, y, 0.5)
    d.norms(x)
    d.unpack(x, y)
    print("ok", p, s, N, B, m, flush=True)
